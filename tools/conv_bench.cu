// Development tool (not product): drives the trunk conv kernels of conv_tc.cu directly on synthetic
// planes, times every layer of the simple net in both kernel modes with CUDA events and, when built
// with -DAP_CONV_TRACE, dumps per-role clock64 timestamps of the first tiles of cluster/CTA 0.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 [-DAP_CONV_TRACE] -o tools/bin/conv_bench tools/conv_bench.cu
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../alphapig_b200/csrc/conv_tc.cu"

int ap_fail(ap_engine* e, int code, const std::string& msg) {
  if (e) e->err = msg;
  return code;
}

int main(int argc, char** argv) {
  const int G = argc > 1 ? atoi(argv[1]) : 4096;
  const int reps = argc > 2 ? atoi(argv[2]) : 5;
  ap_engine e;
  cudaStreamCreate(&e.stream);
  NetState n;
  n.W = n.H = 15;
  n.S = 225;
  n.bcap = G;
  n.mpad = 2ll * NET_PAD_ROWS + (long long)G * NET_TILE_ROWS;
  cudaDeviceGetAttribute(&n.sm_count, cudaDevAttrMultiProcessorCount, 0);
  cudaMalloc(&n.feat, (size_t)2 * n.mpad * 16);
  cudaMemset(n.feat, 0, (size_t)2 * n.mpad * 16);
  for (int i = 0; i < 2; ++i) {
    cudaMalloc(&n.act[i], (size_t)32 * n.mpad * 16);
    cudaMemset(n.act[i], 0, (size_t)32 * n.mpad * 16);
  }
  cudaMalloc(&n.d_err, 4);
  cudaMemset(n.d_err, 0, 4);
  if (conv_tc_configure(&e) != AP_OK) {
    printf("configure failed: %s\n", e.err.c_str());
    return 1;
  }
  const int cin[6] = {9, 64, 64, 128, 128, 256}, cout[6] = {64, 64, 128, 128, 256, 256};
  std::vector<ConvLayer> layers(6);
  for (int i = 0; i < 6; ++i) {
    ConvLayer& L = layers[i];
    L = ConvLayer{};
    L.cin = cin[i];
    L.cin_pad = (cin[i] + 15) & ~15;
    L.cout = cout[i];
    L.relu = 1;
    L.in_buf = i == 0 ? -1 : (i - 1) & 1;
    L.out_buf = i & 1;
    L.resid_buf = -1;
    size_t wb = (size_t)9 * L.cin_pad * L.cout * 2;
    cudaMalloc(&L.wimg, wb);
    cudaMalloc(&L.wimg2, wb);
    cudaMemset(L.wimg, 0, wb);
    cudaMemset(L.wimg2, 0, wb);
    cudaMalloc(&L.shift, L.cout * 4);
    cudaMemset(L.shift, 0, L.cout * 4);
  }
#ifdef AP_CONV_TRACE
  long long* d_trace;
  const size_t trace_n = 2 * 4 * 64 * 8;
  cudaMalloc(&d_trace, trace_n * 8);
  g_conv_trace = d_trace;
#endif
  cudaEvent_t ev[8];
  for (auto& x : ev) cudaEventCreate(&x);
  for (int mode = 1; mode <= 2; ++mode) {
    n.conv_mode = mode;
    float acc[6] = {0, 0, 0, 0, 0, 0};
    for (int r = 0; r < reps + 1; ++r) {
      for (int i = 0; i < 6; ++i) {
        cudaEventRecord(ev[i], e.stream);
        int rc = conv_tc_launch(&e, &n, layers[i], G);
        if (rc != AP_OK) {
          printf("launch failed: %s\n", e.err.c_str());
          return 1;
        }
      }
      cudaEventRecord(ev[6], e.stream);
      cudaError_t st = cudaStreamSynchronize(e.stream);
      if (st != cudaSuccess) {
        printf("CUDA error: %s\n", cudaGetErrorString(st));
        return 1;
      }
      if (r == 0) continue;
      for (int i = 0; i < 6; ++i) {
        float ms;
        cudaEventElapsedTime(&ms, ev[i], ev[i + 1]);
        acc[i] += ms;
      }
    }
    float tot = 0;
    printf("mode %d us/layer:", mode);
    for (int i = 0; i < 6; ++i) {
      printf(" %.1f", 1000.f * acc[i] / reps);
      tot += acc[i] / reps;
    }
    printf("  total %.1f us\n", 1000.f * tot);
#ifdef AP_CONV_TRACE
    const int tl = argc > 3 ? atoi(argv[3]) : 1;  // layer to trace
    cudaMemset(d_trace, 0, trace_n * 8);
    conv_tc_launch(&e, &n, layers[tl], G);
    cudaStreamSynchronize(e.stream);
    std::vector<long long> h(trace_n);
    cudaMemcpy(h.data(), d_trace, trace_n * 8, cudaMemcpyDeviceToHost);
    long long t0 = h[0];
    for (size_t k = 0; k < trace_n; ++k)
      if (h[k] && h[k] < t0) t0 = h[k];
    const char* roles[4] = {"producer", "mma", "epilogue", "relay"};
    for (int b = 0; b < 2; ++b)
      for (int role = 0; role < 4; ++role) {
        printf("mode %d layer %d cta %d %s (cycles since first stamp; slots per tile):\n", mode, tl, b, roles[role]);
        for (int it = 0; it < 12; ++it) {
          printf("  tile %2d:", it);
          for (int s = 0; s < 8; ++s) {
            long long v = h[((b * 4 + role) * 64 + it) * 8 + s];
            if (v) printf(" %7lld", v - t0);
            else printf("       -");
          }
          printf("\n");
        }
      }
#endif
  }
  int herr = 0;
  cudaMemcpy(&herr, n.d_err, 4, cudaMemcpyDeviceToHost);
  printf("errflag %d\n", herr);
  return 0;
}
