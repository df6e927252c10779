// Policy/value net on device (net.cu, conv_tc.cu).
#pragma once
#include <map>

#include "ap_common.cuh"

#define NET_PAD_ROWS 32    // zero rows before the first and after the last board
#define NET_TILE_ROWS 256  // one board = 16x16 padded pixel rows (real pixels x<W, y<H; W,H <= 15)
#define NET_SLAB_ROWS 290  // 256 + 17 halo rows either side

struct ConvLayer {
  int cin, cin_pad, cout;
  int relu;
  int in_buf, out_buf, resid_buf;  // activation buffer ids (resid_buf < 0: none); in_buf -1 = feature buffer
  // offsets into the flat fp32 master buffer (-1 = absent)
  long long w, b, gamma, beta, mean, var;
  int fix_gamma;
  // Inception-ResNet variant: 1x1 convs, channel-slice outputs and composite (block-assembled) weight tensors
  int ksz = 3;            // 3 or 1
  int out_coff = 0;       // output channel offset inside the output planes
  int in_coff = 0;        // input channel offset inside the input planes (multiple of 8): a tower reads its slice only
  int cout_store = 0;     // channels stored (0 = cout)
  int force_single = 0;   // always the single-CTA kernel (the slice stores / 1x1 taps live there)
  int virt = 0;           // parameter offsets index NetState::vmaster (assembled by k_build_virtual) instead of master
  float post_scale = 1.f; // folded after BN (block35: net += 0.17 * up)
  // prepared (device)
  __half* wimg = nullptr;  // [nkc][9][KC/8][cout][8]
  __half* wimg2 = nullptr; // CTA-pair image [2][nkc][9][KC/8][cout/2][8]: half r holds output channels [r*cout/2, (r+1)*cout/2)
  __half* wimg4 = nullptr; // cout 256 only: [2 column halves][2 ranks][nkc][9][KC/8][64][8] for k_conv3x3_tc4
  __half* wimg_lo = nullptr;  // split precision only: fp16(w - fp16(w)) in the wimg layout (K chunk 32)
  float* scale = nullptr;  // [cout]
  float* shift = nullptr;  // [cout]
};

struct HeadParams {
  int cfin;  // channels of the trunk output
  long long pw, pb, pgamma, pbeta, pmean, pvar;  // conv3_1_1 (4 ch)
  long long vw, vb, vgamma, vbeta, vmean, vvar;  // conv3_2_1 (2 ch)
  long long fcp_w, fcp_b, fcv_w, fcv_b;
  float* w1x1 = nullptr;   // [6][cfin] scale-folded
  float* b1x1 = nullptr;   // [6]
  float* fcpT = nullptr;   // [4S][232] transposed + zero padded, k index in tensor-pixel order
  float* fcp_bias = nullptr;
  float* fcv = nullptr;    // [2S]
  float* fcv_bias = nullptr;
};

// host copy of the scale-folded 1x1 head-conv weights, passed BY VALUE to the fused-head conv kernels
struct NetHeadW {
  float w[6 * 256];  // [6][cfin] (cfin <= 256): 4 policy + 2 value output channels
  float b[6];
};

struct NetState {
  int arch = 0, n_blocks = 0, n_filter = 0;
  int W = 0, H = 0, S = 0;
  int bcap = 0;         // boards the activation buffers hold
  long long mpad = 0;   // rows per channel-group plane
  std::vector<std::string> names;
  std::vector<long long> offsets, numels;
  std::map<std::string, int> index;
  float* master = nullptr;
  long long master_numel = 0;
  // Inception-ResNet variant: dense "virtual" conv parameters assembled from several named convs
  struct VCopy { long long dst_w, dst_b, dst_beta, dst_mean, dst_var, src_w, src_b, src_beta, src_mean, src_var;
                 int cout, cin, cin_v, taps, oo, io; };
  float* vmaster = nullptr;
  long long vmaster_numel = 0;
  std::vector<long long> vvar_ranges;  // (offset, count) pairs of virtual `var` arrays: reset to 1 before assembling
  std::vector<VCopy> vcopies;
  std::vector<ConvLayer> trunk;
  HeadParams head;
  __half* feat = nullptr;      // [2][mpad][8]
  __half* act[4] = {nullptr, nullptr, nullptr, nullptr};  // [32][mpad][8]
  // split precision (AP_NET_SPLIT): every activation is a hi + lo fp16 pair (22 significant bits), every weight too,
  // and the conv kernels issue hi*hi + lo*hi + hi*lo: what the 10-block residual net needs to stay within 1e-3
  int split = 0;
  __half* feat_lo = nullptr;  // all zero (features are exactly 0/1)
  __half* act_lo[3] = {nullptr, nullptr, nullptr};
  int final_buf = 0;
  float* hbuf = nullptr;  // [bcap][6][S] fp32 outputs of the two 1x1 head convs (legacy fp32 FC path only)
  // tensor-core FC heads (heads_tc.cu): split-fp16 operands
  __half* fc_a = nullptr;    // [2 (hi,lo)][fc_kp/8][fc_rows][8]  head-conv outputs, k = o*S + pixel
  __half* fc_w = nullptr;    // [2][fc_kp/8][fc_np][8]
  float* fc_bias = nullptr;  // [fc_np]
  float* fc_partial = nullptr;  // [fc_ksplit][fc_rows][fc_np] raw partial logits of the split-K FC
  int fc_ksplit = 4;
  long long fc_rows = 0;     // bcap rounded up to 128
  int fc_kp = 0, fc_np = 0;  // 6S rounded up to the K stage; S+1 rounded up to 16
  int head_mode = 2;         // 0 = fp32 CUDA-core FC (legacy), 1 = k_head_conv + tensor-core FC, 2 = head convs fused into
                             // the last trunk epilogue + tensor-core FC; AP_HEAD_MODE overrides for A/B timing
  // fp32 CUDA-core reference path
  float* ref_a = nullptr;  // [bcap_ref][256][S]
  float* ref_b = nullptr;
  float* ref_c = nullptr;
  float* ref_in = nullptr;  // [bcap_ref][9][S]
  int bcap_ref = 0;
  int sm_count = 148;
  int conv_mode = 0;  // 0 = per-layer choice, 1 = single-CTA kernel, 2 = CTA-pair kernel (cta_group::2); AP_CONV_MODE overrides for A/B timing
  int conv4 = 2;      // 256-channel layers on the two-boards-per-pair kernel: 0 off, 1 plain layers, 2 also the
                      // fused-head layer (AP_CONV4 for A/B timing)
  int conv4_128 = 0;  // 128-channel layers on that kernel too: 1 = cin 64 (conv3), 2 = all (AP_CONV4_128, A/B timing)
  NetHeadW head_w;
  int pair_cin64 = 1;   // layers with 64 input and 128 output channels (conv3) on the CTA-pair kernel too: 0.116 -> 0.108 ms per
                        // lock-step on B200 (AP_PAIR_CIN64=0 for A/B)
  int front_fused = 1;  // conv1 + conv2 of the 6-conv net as one kernel (front_tc.cu); AP_FRONT_FUSED=0 runs the two layers
                        // separately (A/B timing and the bit-equality test)
  int head_pair = 1;  // run the fused-head layer on the CTA-pair kernel (measured faster: its double-buffered TMEM hides
                      // the longer epilogue); AP_HEAD_PAIR=0 selects the single-CTA kernel for A/B timing
  int* d_err = nullptr;
  std::vector<void*> allocs;
};

// conv_tc.cu
int conv_tc_launch(ap_engine* e, NetState* n, const ConvLayer& L, int n_boards, const int* n_boards_dev = nullptr,
                   bool head = false);
bool conv_tc_head_supported(const ConvLayer& L);
bool conv_tc_split_supported(const ConvLayer& L);
int conv_tc_kc(const ConvLayer& L, int split);
bool conv_tc_supported(int cin_pad, int cout);
int conv_tc_smem_bytes(const ConvLayer& L, int* out_nb);
int conv_tc_configure(ap_engine* e);
// front_tc.cu
bool front_tc_supported(const NetState* n);
int front_tc_configure(ap_engine* e);
int front_tc_launch(ap_engine* e, NetState* n, int n_boards, const int* n_boards_dev);
// heads_tc.cu
void fc_tc_dims(int S, int* kp, int* np);
int fc_tc_configure(ap_engine* e, NetState* n);
int fc_tc_prep(ap_engine* e, NetState* n);
int fc_tc_launch(ap_engine* e, NetState* n, int nb, float* d_probs, float* d_values, const int* nb_dev = nullptr,
                 bool skip_finish = false);
