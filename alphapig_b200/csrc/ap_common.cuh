// Shared definitions for libalphapig_b200 (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <vector>

#include "../../include/alphapig_b200.h"

#define AP_MAX_S 256
#define AP_ROWS 16
#define AP_FULL 0xffffffffu

// Per-game board meta (16 B).  hist[0] is the most recent move, -1 = none.
struct BoardMeta {
  int16_t hist[4];
  int16_t n_stones;
  int16_t last_move;
  int8_t cur;    // player to move: 1 or 2
  int8_t start;  // start_player index given to init_board
  int16_t pad;
};

// Geometry + constants passed by value to kernels.
struct Geo {
  int W, H, S, n_in_row;
  int G;
  int cap;  // node capacity per game
  double c_puct;
};

// SoA node pools: index = g*cap + i.  32 B / node.
struct Pools {
  double* P;
  double* Q;
  int32_t* N;
  int32_t* child_start;  // -1 = leaf
  int32_t* parent;       // -1 = root
  uint16_t* child_count;
  int16_t* move;
  int32_t* alloc;  // [G] nodes in use
};

// Per-game leaf record written by select, read by features / expand / backup.
struct Leaves {
  uint32_t* rows;   // [G][16]
  BoardMeta* meta;  // [G]
  int32_t* node;    // [G]
  int8_t* terminal; // [G] 0/1
  int8_t* winner;   // [G] winner at the leaf (1,2,-1) when terminal
  int32_t* depth;   // [G]
  int16_t* path;    // [G][S]
  // compaction of the non-terminal leaves for the device-net path (terminal leaves never reach the net)
  int32_t* slot;          // [G] net tile of game g, -1 = terminal leaf
  int32_t* game_of_slot;  // [G]
  int32_t* n_eval;        // [1] non-terminal leaves of the last compaction
  // ap_search_set_active: games with active[g] == 0 are skipped by every search kernel (finished arena games)
  const uint8_t* active;  // [G]
};

struct NetState;     // net.cu
struct ReplayState;  // replay.cu
struct TrajState;    // traj.cu

struct ap_engine {
  ap_config cfg;
  Geo geo;
  cudaStream_t stream = nullptr;
  std::string err;
  uint64_t launches = 0;
  uint64_t bytes = 0;
  std::vector<void*> allocs;
  // boards
  uint32_t* rows = nullptr;  // [G][16]  lo16 = player 1 stones of board row h, hi16 = player 2
  BoardMeta* meta = nullptr; // [G]
  Pools pools{};
  Leaves leaves{};
  int32_t* errflag = nullptr;     // [G] device error codes (sticky until read)
  unsigned long long* stats = nullptr;  // [8] device counters
  // scratch for compaction
  void* scratch = nullptr;
  int scratch_slots = 0;
  // staging
  void* d_stage = nullptr;
  size_t stage_bytes = 0;
  void* h_stage = nullptr;  // pinned
  size_t h_stage_bytes = 0;
  int32_t* d_ids = nullptr;  // [G]
  // net
  NetState* net = nullptr;
  ReplayState* replay = nullptr;
  TrajState* traj = nullptr;
  float* d_probs = nullptr;   // [G][S] fp32
  float* d_values = nullptr;  // [G]
  // ap_pure_run leaves lazily materialised trees (only the root's child block is complete, see rollout.cu); the
  // reference's mcts_pure never reuses a tree (get_action ends with update_with_move(-1), mcts_pure.py:196-203), so
  // the next tree operation other than reading the root starts from fresh roots
  bool pure_tree = false;
  // node pools.  cap_auto: the capacity was picked by the library (cfg.node_capacity <= 0) and GROWS on demand - the
  // reference's trees are unbounded Python objects and a re-rooted subtree keeps accumulating visits over the plies of
  // a game (root N -> n_playout / (1 - share of the chosen child)), so no fixed bound is safe for self-play with tree
  // reuse.  An explicit capacity is a hard limit (AP_ERR_POOL_EXHAUSTED when a game runs over it).
  bool cap_auto = false;
  uint64_t pool_generation = 0;  // bumped when the pools move (captured graphs hold the old pointers)
  int32_t* d_max_alloc = nullptr;
  // small batches (interactive play: one game) are launch-latency bound: the n_playout lock-steps of ap_search_run are
  // captured once into a CUDA graph and replayed; rebuilt when n_playout or the prepared weights change
  // opt-in multi-leaf mode (ap_search_run_vl): leaf records for G * vl_kstride leaves, per-node in-flight visit counts
  Leaves leaves_vl{};
  int vl_kstride = 0;
  int32_t* vn = nullptr;  // [G][cap] virtual visits (all zero between lock-steps)
  int vn_cap = 0;
  int32_t* vl_remain = nullptr;  // [G]
  int32_t* vl_issued = nullptr;  // [G]
  cudaGraphExec_t vl_graph = nullptr;
  int vl_graph_playouts = 0, vl_graph_k = 0;
  uint64_t vl_graph_gen = 0, vl_graph_launches = 0;
  cudaGraphExec_t run_graph = nullptr;
  int run_graph_playouts = 0;
  uint64_t run_graph_gen = 0, run_graph_launches = 0;
  uint64_t net_generation = 0;  // bumped by every weight preparation (kernel-parameter copies of the head weights)
  float last_total_ms = 0.f, last_net_ms = 0.f;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  // ap_search_run keeps at most two chunks of lock-steps queued on the stream (the host waits for the event of the
  // chunk before last): a bottomless launch queue would block every other thread of the process that wants to launch
  // on this GPU - the trainer of alphapig_b200/loop.py - behind thousands of search kernels
  cudaEvent_t chunk_ev[3] = {nullptr, nullptr, nullptr};
  // optional per-phase timing of ap_search_run (ap_search_profile): events after every phase of every lock-step
  int profile = 0;
  std::vector<cudaEvent_t> prof_events;
  std::vector<float> prof_ms;  // accumulated ms per phase of the last ap_search_run
  int prof_cursor = 0;
};

// every C-ABI entry point: validate the handle and make its device current for the calling thread (worker threads
// of the Python drivers start on device 0)
#define AP_ENTER(e)                           \
  do {                                        \
    if (!(e)) return AP_ERR_BAD_HANDLE;       \
    cudaSetDevice((e)->cfg.device);           \
  } while (0)

#define AP_CUDA(e, call)                                                                  \
  do {                                                                                    \
    cudaError_t _st = (call);                                                             \
    if (_st != cudaSuccess) {                                                             \
      (e)->err = std::string(#call) + ": " + cudaGetErrorString(_st);                     \
      return AP_ERR_CUDA;                                                                 \
    }                                                                                     \
  } while (0)

#define AP_LAUNCH_CHECK(e)                                                                \
  do {                                                                                    \
    (e)->launches++;                                                                      \
    cudaError_t _st = cudaGetLastError();                                                 \
    if (_st != cudaSuccess) {                                                             \
      (e)->err = std::string("kernel launch: ") + cudaGetErrorString(_st) + " at " +      \
                 __FILE__ + ":" + std::to_string(__LINE__);                               \
      return AP_ERR_CUDA;                                                                 \
    }                                                                                     \
  } while (0)

// helpers implemented in engine.cu
int ap_fail(ap_engine* e, int code, const std::string& msg);
int ap_stage(ap_engine* e, size_t dbytes, size_t hbytes);
int ap_ids(ap_engine* e, const int32_t* game_ids, int32_t n);  // uploads ids (or iota) into e->d_ids

// replay.cu
void replay_destroy(ap_engine* e);
// traj.cu
void traj_destroy(ap_engine* e);
int traj_append_pick(ap_engine* e, const float* d_pi);
int traj_record_width(int S);
// net.cu
int net_destroy(ap_engine* e);
int net_forward_leaves(ap_engine* e, int precise, bool compact = false, bool compacted_by_select = false);
int net_emit_features_launch(ap_engine* e, bool compact = false);
int net_check_err(ap_engine* e);
int net_phase_count(ap_engine* e);
bool net_can_compact(ap_engine* e);
int net_board_capacity(ap_engine* e);
int net_run_compacted(ap_engine* e, int nb_max, const int32_t* nb_dev);
void net_feature_planes(ap_engine* e, __half** feat, long long* mpad);
void net_fc_finish_args(ap_engine* e, const float** partial, const float** bias, long long* rows, int* np, int* ksplit);
void prof_mark(ap_engine* e);
