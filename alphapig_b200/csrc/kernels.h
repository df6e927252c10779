// Host-side launch wrappers (each enqueues on e->stream; caller does AP_LAUNCH_CHECK).
#pragma once
#include "ap_common.cuh"

// boards.cu
void launch_boards_reset(ap_engine* e, int n, const int32_t* d_start);
void launch_boards_do_move(ap_engine* e, int n, const int32_t* d_moves, int32_t* d_status);
void launch_boards_status(ap_engine* e, const uint32_t* rows, const BoardMeta* meta, const int32_t* d_ids, int n,
                          uint8_t* d_end, int8_t* d_winner);
void launch_boards_legal(ap_engine* e, const int32_t* d_ids, int n, uint32_t* d_mask);
void launch_boards_features(ap_engine* e, const uint32_t* rows, const BoardMeta* meta, const int32_t* d_ids, int n,
                            float* d_out);
void launch_pack_bits(ap_engine* e, const float* d_f, int n, int nbits, uint8_t* d_out);
void launch_boards_export(ap_engine* e, const uint32_t* rows, const BoardMeta* meta, const int32_t* d_ids, int n,
                          int8_t* d_cells, int32_t* d_meta);
void launch_boards_import(ap_engine* e, int n, const int8_t* d_cells, const int32_t* d_meta);

// tree.cu
void launch_tree_reset_all(ap_engine* e);
void launch_select(ap_engine* e, bool compact = false);
void launch_expand_backup(ap_engine* e, const int32_t* d_counts, const int16_t* d_acts, const double* d_pri64,
                          const double* d_val64, const float* d_pri32, const float* d_val32,
                          const int32_t* d_slot = nullptr, bool fuse_fc_finish = false);
void launch_compact_leaves(ap_engine* e);
void launch_select_vl(ap_engine* e, const Leaves& lv, int32_t* vn, int k, int kstride, int32_t* remain, int32_t* issued);
void launch_expand_backup_vl(ap_engine* e, const Leaves& lv, int32_t* vn, int kstride, const int32_t* issued);
void launch_selfplay_pick(ap_engine* e, double temp, double eps, double alpha, uint64_t seed, uint32_t ply, int32_t* d_move,
                          float* d_pi, double* d_noise);
void launch_advance(ap_engine* e, int n, const int32_t* d_moves);
void launch_root(ap_engine* e, const int32_t* d_ids, int n, int32_t* d_count, int16_t* d_acts, int32_t* d_visits,
                 double* d_q, int32_t* d_rootn);
void launch_root_probs(ap_engine* e, double temp, double* d_out);
size_t scratch_bytes_per_slot(int cap);
void launch_max_alloc(ap_engine* e, int32_t* d_out);
void launch_pool_copy(ap_engine* e, const Pools& from, int from_cap, const Pools& to, int to_cap);

// rollout.cu
void launch_pure_run(ap_engine* e, int n_playout, uint64_t seed, int mode, int32_t* d_move);
void launch_rollout_eval(ap_engine* e, uint64_t seed, int impl, const uint32_t* d_keys, int8_t* d_value, int16_t* d_plies);
void launch_rollout_hash(ap_engine* e, int8_t* d_value);
