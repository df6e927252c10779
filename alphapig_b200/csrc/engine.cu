// C-ABI of libalphapig_b200.so: handle lifecycle, staging, and the host side of every entry point.
#include <string.h>

#include <algorithm>

#include "kernels.h"
#include "net.h"

int ap_fail(ap_engine* e, int code, const std::string& msg) {
  if (e) e->err = msg;
  return code;
}

static int dev_alloc(ap_engine* e, void** p, size_t bytes, bool zero = true) {
  AP_CUDA(e, cudaMalloc(p, bytes));
  e->allocs.push_back(*p);
  e->bytes += bytes;
  if (zero) AP_CUDA(e, cudaMemsetAsync(*p, 0, bytes, e->stream));
  return AP_OK;
}
#define AP_TRY(x)            \
  do {                       \
    int _r = (x);            \
    if (_r != AP_OK) return _r; \
  } while (0)

// grow-only device + pinned host staging buffers
int ap_stage(ap_engine* e, size_t dbytes, size_t hbytes) {
  if (dbytes > e->stage_bytes) {
    if (e->d_stage) cudaFree(e->d_stage);
    e->d_stage = nullptr;
    size_t nb = std::max(dbytes, e->stage_bytes * 2);
    AP_CUDA(e, cudaMalloc(&e->d_stage, nb));
    e->stage_bytes = nb;
  }
  if (hbytes > e->h_stage_bytes) {
    if (e->h_stage) cudaFreeHost(e->h_stage);
    e->h_stage = nullptr;
    size_t nb = std::max(hbytes, e->h_stage_bytes * 2);
    AP_CUDA(e, cudaMallocHost(&e->h_stage, nb));
    e->h_stage_bytes = nb;
  }
  return AP_OK;
}

int ap_ids(ap_engine* e, const int32_t* game_ids, int32_t n) {
  if (n < 0 || n > e->geo.G) return ap_fail(e, AP_ERR_BAD_ARG, "n out of range");
  std::vector<int32_t> ids(n);
  for (int i = 0; i < n; ++i) {
    ids[i] = game_ids ? game_ids[i] : i;
    if (ids[i] < 0 || ids[i] >= e->geo.G) return ap_fail(e, AP_ERR_BAD_ARG, "game id out of range");
  }
  AP_CUDA(e, cudaMemcpyAsync(e->d_ids, ids.data(), sizeof(int32_t) * n, cudaMemcpyHostToDevice, e->stream));
  AP_CUDA(e, cudaStreamSynchronize(e->stream));  // ids vector dies at return
  return AP_OK;
}

static int h2d(ap_engine* e, void* d, const void* h, size_t bytes) {
  AP_CUDA(e, cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice, e->stream));
  return AP_OK;
}
static int d2h_sync(ap_engine* e, void* h, const void* d, size_t bytes) {
  AP_CUDA(e, cudaMemcpyAsync(h, d, bytes, cudaMemcpyDeviceToHost, e->stream));
  AP_CUDA(e, cudaStreamSynchronize(e->stream));
  return AP_OK;
}

static int check_errflags(ap_engine* e) {
  // sticky per-game device error codes -> first one reported
  std::vector<int32_t> f(e->geo.G);
  AP_TRY(d2h_sync(e, f.data(), e->errflag, sizeof(int32_t) * e->geo.G));
  for (int g = 0; g < e->geo.G; ++g)
    if (f[g] != 0) {
      int code = f[g];
      cudaMemsetAsync(e->errflag, 0, sizeof(int32_t) * e->geo.G, e->stream);
      return ap_fail(e, code, "game " + std::to_string(g) + ": " +
                                  (code == AP_ERR_POOL_EXHAUSTED ? "node pool exhausted (raise node_capacity)"
                                                                 : "device error"));
    }
  return AP_OK;
}

static void dev_free(ap_engine* e, void* p, size_t bytes) {
  if (!p) return;
  auto it = std::find(e->allocs.begin(), e->allocs.end(), p);
  if (it != e->allocs.end()) e->allocs.erase(it);
  cudaFree(p);
  e->bytes -= bytes;
}
static size_t pool_bytes(size_t G, size_t cap) { return G * cap * 32; }

static int alloc_pools(ap_engine* e, int cap, Pools* pl, void** scratch) {
  const size_t G = e->geo.G, N = G * (size_t)cap;
  AP_TRY(dev_alloc(e, (void**)&pl->P, N * 8));
  AP_TRY(dev_alloc(e, (void**)&pl->Q, N * 8));
  AP_TRY(dev_alloc(e, (void**)&pl->N, N * 4));
  AP_TRY(dev_alloc(e, (void**)&pl->child_start, N * 4));
  AP_TRY(dev_alloc(e, (void**)&pl->parent, N * 4));
  AP_TRY(dev_alloc(e, (void**)&pl->child_count, N * 2));
  AP_TRY(dev_alloc(e, (void**)&pl->move, N * 2));
  AP_TRY(dev_alloc(e, (void**)&pl->alloc, G * 4));
  AP_TRY(dev_alloc(e, scratch, (size_t)e->scratch_slots * scratch_bytes_per_slot(cap)));
  return AP_OK;
}
static void free_pools(ap_engine* e, int cap, Pools* pl, void* scratch) {
  const size_t G = e->geo.G, N = G * (size_t)cap;
  dev_free(e, pl->P, N * 8);
  dev_free(e, pl->Q, N * 8);
  dev_free(e, pl->N, N * 4);
  dev_free(e, pl->child_start, N * 4);
  dev_free(e, pl->parent, N * 4);
  dev_free(e, pl->child_count, N * 2);
  dev_free(e, pl->move, N * 2);
  dev_free(e, pl->alloc, G * 4);
  dev_free(e, scratch, (size_t)e->scratch_slots * scratch_bytes_per_slot(cap));
  *pl = Pools{};
}

// Make room for `need_free` more nodes in every game's pool.  Library-chosen capacities grow (the reference's trees
// are unbounded); an explicit node_capacity is a hard limit and a game that runs over it reports
// AP_ERR_POOL_EXHAUSTED from the expanding kernel.  One 4-byte D2H per call.
static int ensure_pool(ap_engine* e, long long need_free) {
  if (!e->cap_auto) return AP_OK;
  int32_t mx = 0;
  launch_max_alloc(e, e->d_max_alloc);
  AP_LAUNCH_CHECK(e);
  AP_TRY(d2h_sync(e, &mx, e->d_max_alloc, 4));
  const long long need = (long long)mx + need_free;
  if (need <= e->geo.cap) return AP_OK;
  long long want = std::max(need, (long long)e->geo.cap * 3 / 2);
  want = (want + 3) & ~3ll;
  const size_t G = e->geo.G;
  size_t fr = 0, tot = 0;
  cudaMemGetInfo(&fr, &tot);
  auto cost = [&](long long c) { return pool_bytes(G, (size_t)c) + (size_t)e->scratch_slots * scratch_bytes_per_slot((int)c); };
  const size_t margin = (size_t)1 << 30;
  if (cost(want) + margin > fr) want = (need + 3) & ~3ll;  // no head-room: take exactly what this search needs
  if (cost(want) + margin > fr || want > (1ll << 26))
    return ap_fail(e, AP_ERR_POOL_EXHAUSTED,
                   "node pools cannot grow to " + std::to_string(want) + " nodes per game (" +
                       std::to_string(cost(want) >> 20) + " MiB needed, " + std::to_string(fr >> 20) + " MiB free)");
  Pools np{};
  void* nscratch = nullptr;
  const int rc = alloc_pools(e, (int)want, &np, &nscratch);
  if (rc != AP_OK) {
    free_pools(e, (int)want, &np, nscratch);
    return ap_fail(e, AP_ERR_POOL_EXHAUSTED, "node pools cannot grow: " + e->err);
  }
  launch_pool_copy(e, e->pools, e->geo.cap, np, (int)want);
  AP_LAUNCH_CHECK(e);
  AP_CUDA(e, cudaStreamSynchronize(e->stream));
  free_pools(e, e->geo.cap, &e->pools, e->scratch);
  e->pools = np;
  e->scratch = nscratch;
  e->geo.cap = (int)want;
  e->pool_generation++;
  return AP_OK;
}

void prof_mark(ap_engine* e) {
  if (!e->profile) return;
  if ((size_t)e->prof_cursor < e->prof_events.size()) cudaEventRecord(e->prof_events[e->prof_cursor++], e->stream);
}

extern "C" {

const char* ap_version(void) { return "alphapig_b200 0.1 (sm_100a)"; }

const char* ap_last_error(const ap_engine* e) { return e ? e->err.c_str() : "bad handle"; }

int ap_engine_create(const ap_config* cfg, ap_engine** out) {
  if (!cfg || !out) return AP_ERR_BAD_ARG;
  *out = nullptr;
  // game.py:36-38: width/height must be >= n_in_row
  if (cfg->width < cfg->n_in_row || cfg->height < cfg->n_in_row || cfg->width > 16 || cfg->height > 16 ||
      cfg->n_in_row < 2 || cfg->n_games < 1)
    return AP_ERR_BAD_ARG;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || cfg->device < 0 || cfg->device >= ndev) return AP_ERR_CUDA;
  ap_engine* e = new ap_engine();
  e->cfg = *cfg;
  if (cudaSetDevice(cfg->device) != cudaSuccess) {
    delete e;
    return AP_ERR_CUDA;
  }
  Geo& g = e->geo;
  g.W = cfg->width;
  g.H = cfg->height;
  g.S = g.W * g.H;
  g.n_in_row = cfg->n_in_row;
  g.G = cfg->n_games;
  g.c_puct = cfg->c_puct;
  int cap = cfg->node_capacity;
  e->cap_auto = cap <= 0;
  if (cap <= 0) {
    // every playout adds at most S children; a re-rooted subtree usually retains about one search's worth - when a
    // game keeps more (ensure_pool) the pools grow
    long long want = 2ll * std::max(cfg->n_playout_hint, 1) * g.S + g.S + 2;
    cap = (int)std::min<long long>(want, 1 << 20);
  }
  g.cap = (cap + 3) & ~3;  // SoA pools and the compaction scratch slots keep every array 8-byte aligned
  int rc = AP_OK;
  auto fail = [&](int code) {
    e->err += " (engine_create)";
    for (void* p : e->allocs) cudaFree(p);
    if (e->stream) cudaStreamDestroy(e->stream);
    static std::string last;
    last = e->err;
    delete e;
    return code;
  };
  {
    int lo = 0, hi = 0;  // numerically lower = higher priority
    cudaDeviceGetStreamPriorityRange(&lo, &hi);
    const int prio = (cfg->flags & AP_FLAG_HIGH_PRIORITY_STREAM) ? hi : lo;
    if (cudaStreamCreateWithPriority(&e->stream, cudaStreamNonBlocking, prio) != cudaSuccess) return fail(AP_ERR_CUDA);
  }
  const size_t G = g.G;
#define ALLOC(ptr, bytes) \
  if ((rc = dev_alloc(e, (void**)&(ptr), (bytes))) != AP_OK) return fail(rc)
  ALLOC(e->rows, G * AP_ROWS * 4);
  ALLOC(e->meta, G * sizeof(BoardMeta));
  e->scratch_slots = (int)std::min<size_t>(G, 296);
  if ((rc = alloc_pools(e, g.cap, &e->pools, &e->scratch)) != AP_OK) return fail(rc);
  ALLOC(e->d_max_alloc, 4);
  ALLOC(e->leaves.rows, G * AP_ROWS * 4);
  ALLOC(e->leaves.meta, G * sizeof(BoardMeta));
  ALLOC(e->leaves.node, G * 4);
  ALLOC(e->leaves.terminal, G);
  ALLOC(e->leaves.winner, G);
  ALLOC(e->leaves.depth, G * 4);
  ALLOC(e->leaves.path, G * (size_t)g.S * 2);
  ALLOC(e->leaves.slot, G * 4);
  ALLOC(e->leaves.game_of_slot, G * 4);
  ALLOC(e->leaves.n_eval, 4);
  {
    uint8_t* act = nullptr;
    ALLOC(act, G);
    cudaMemsetAsync(act, 1, G, e->stream);
    e->leaves.active = act;
  }
  ALLOC(e->errflag, G * 4);
  ALLOC(e->stats, 8 * 8);
  ALLOC(e->d_ids, G * 4);
  ALLOC(e->d_probs, G * (size_t)g.S * 4);
  ALLOC(e->d_values, G * 4);
#undef ALLOC
  cudaEventCreate(&e->ev0);
  cudaEventCreate(&e->ev1);
  // all boards empty with player 1 to move, all trees a fresh root
  std::vector<int32_t> ids(G);
  for (size_t i = 0; i < G; ++i) ids[i] = (int32_t)i;
  cudaMemcpyAsync(e->d_ids, ids.data(), G * 4, cudaMemcpyHostToDevice, e->stream);
  launch_boards_reset(e, g.G, nullptr);
  launch_tree_reset_all(e);
  e->launches += 2;
  if (cudaStreamSynchronize(e->stream) != cudaSuccess || cudaGetLastError() != cudaSuccess) {
    e->err = "initial reset kernels failed (is this an sm_100a device?)";
    return fail(AP_ERR_CUDA);
  }
  *out = e;
  return AP_OK;
}

int ap_engine_node_capacity(const ap_engine* e, int32_t* out_nodes) {
  if (!e || !out_nodes) return AP_ERR_BAD_HANDLE;
  *out_nodes = e->geo.cap;
  return AP_OK;
}

int ap_engine_destroy(ap_engine* e) {
  if (e && e->run_graph) {
    cudaGraphExecDestroy(e->run_graph);
    e->run_graph = nullptr;
  }
  if (e && e->vl_graph) {
    cudaGraphExecDestroy(e->vl_graph);
    e->vl_graph = nullptr;
  }
  AP_ENTER(e);
  cudaSetDevice(e->cfg.device);
  cudaStreamSynchronize(e->stream);
  net_destroy(e);
  replay_destroy(e);
  traj_destroy(e);
  for (void* p : e->allocs) cudaFree(p);
  if (e->d_stage) cudaFree(e->d_stage);
  if (e->h_stage) cudaFreeHost(e->h_stage);
  for (cudaEvent_t ev : e->prof_events) cudaEventDestroy(ev);
  for (cudaEvent_t ev : e->chunk_ev)
    if (ev) cudaEventDestroy(ev);
  if (e->ev0) cudaEventDestroy(e->ev0);
  if (e->ev1) cudaEventDestroy(e->ev1);
  cudaStreamDestroy(e->stream);
  delete e;
  return AP_OK;
}

int ap_sync(ap_engine* e) {
  AP_ENTER(e);
  AP_CUDA(e, cudaStreamSynchronize(e->stream));
  return AP_OK;
}

int ap_engine_memory(const ap_engine* e, uint64_t* out_bytes) {
  if (!e || !out_bytes) return AP_ERR_BAD_HANDLE;
  *out_bytes = e->bytes;
  return AP_OK;
}

int ap_launch_count(const ap_engine* e, uint64_t* out) {
  if (!e || !out) return AP_ERR_BAD_HANDLE;
  *out = e->launches;
  return AP_OK;
}

// ---- boards -------------------------------------------------------------------------------

int ap_boards_reset(ap_engine* e, const int32_t* game_ids, int32_t n, const int32_t* start_player) {
  AP_ENTER(e);
  AP_TRY(ap_ids(e, game_ids, n));
  int32_t* d_start = nullptr;
  if (start_player) {
    for (int i = 0; i < n; ++i)
      if (start_player[i] != 0 && start_player[i] != 1)
        return ap_fail(e, AP_ERR_BAD_ARG, "start_player should be either 0 or 1");  // game.py:206-208
    AP_TRY(ap_stage(e, sizeof(int32_t) * n, 0));
    d_start = (int32_t*)e->d_stage;
    AP_TRY(h2d(e, d_start, start_player, sizeof(int32_t) * n));
  }
  launch_boards_reset(e, n, d_start);
  AP_LAUNCH_CHECK(e);
  return ap_sync(e);
}

int ap_boards_do_move(ap_engine* e, const int32_t* game_ids, const int32_t* moves, int32_t n, int32_t* out_status) {
  AP_ENTER(e);
  AP_TRY(ap_ids(e, game_ids, n));
  AP_TRY(ap_stage(e, sizeof(int32_t) * 2 * n, 0));
  int32_t* d_moves = (int32_t*)e->d_stage;
  int32_t* d_status = d_moves + n;
  AP_TRY(h2d(e, d_moves, moves, sizeof(int32_t) * n));
  launch_boards_do_move(e, n, d_moves, d_status);
  AP_LAUNCH_CHECK(e);
  std::vector<int32_t> st(n);
  AP_TRY(d2h_sync(e, st.data(), d_status, sizeof(int32_t) * n));
  int rc = AP_OK;
  for (int i = 0; i < n; ++i) {
    if (out_status) out_status[i] = st[i];
    if (st[i] != AP_OK && rc == AP_OK) {
      rc = st[i];
      e->err = "illegal move " + std::to_string(moves[i]) + " for game " + std::to_string(game_ids ? game_ids[i] : i);
    }
  }
  return rc;
}

int ap_boards_status(ap_engine* e, const int32_t* game_ids, int32_t n, uint8_t* out_end, int8_t* out_winner) {
  AP_ENTER(e);
  AP_TRY(ap_ids(e, game_ids, n));
  AP_TRY(ap_stage(e, 2 * (size_t)n, 0));
  uint8_t* d_end = (uint8_t*)e->d_stage;
  int8_t* d_win = (int8_t*)e->d_stage + n;
  launch_boards_status(e, e->rows, e->meta, e->d_ids, n, d_end, d_win);
  AP_LAUNCH_CHECK(e);
  AP_CUDA(e, cudaMemcpyAsync(out_end, d_end, n, cudaMemcpyDeviceToHost, e->stream));
  return d2h_sync(e, out_winner, d_win, n);
}

int ap_boards_legal(ap_engine* e, const int32_t* game_ids, int32_t n, uint32_t* out_mask) {
  AP_ENTER(e);
  AP_TRY(ap_ids(e, game_ids, n));
  AP_TRY(ap_stage(e, 32 * (size_t)n, 0));
  launch_boards_legal(e, e->d_ids, n, (uint32_t*)e->d_stage);
  AP_LAUNCH_CHECK(e);
  return d2h_sync(e, out_mask, e->d_stage, 32 * (size_t)n);
}

int ap_boards_features(ap_engine* e, const int32_t* game_ids, int32_t n, float* out) {
  AP_ENTER(e);
  AP_TRY(ap_ids(e, game_ids, n));
  size_t bytes = (size_t)n * 9 * e->geo.S * 4;
  AP_TRY(ap_stage(e, bytes, 0));
  launch_boards_features(e, e->rows, e->meta, e->d_ids, n, (float*)e->d_stage);
  e->launches++;
  AP_LAUNCH_CHECK(e);
  return d2h_sync(e, out, e->d_stage, bytes);
}

// Board.current_state() bit-packed (np.packbits of the 9xWxH planes): what the replay ring stores
int ap_boards_features_packed(ap_engine* e, const int32_t* game_ids, int32_t n, uint8_t* out) {
  AP_ENTER(e);
  AP_TRY(ap_ids(e, game_ids, n));
  const int nbits = 9 * e->geo.S, sb = (nbits + 7) / 8;
  const size_t fb = ((size_t)n * nbits * 4 + 15) & ~(size_t)15;
  AP_TRY(ap_stage(e, fb + (size_t)n * sb, 0));
  launch_boards_features(e, e->rows, e->meta, e->d_ids, n, (float*)e->d_stage);
  e->launches++;
  AP_LAUNCH_CHECK(e);
  uint8_t* d_out = (uint8_t*)e->d_stage + fb;
  launch_pack_bits(e, (const float*)e->d_stage, n, nbits, d_out);
  AP_LAUNCH_CHECK(e);
  return d2h_sync(e, out, d_out, (size_t)n * sb);
}

static int export_common(ap_engine* e, const uint32_t* rows, const BoardMeta* meta, const int32_t* d_ids, int n,
                         int8_t* out_cells, int32_t* out_meta) {
  size_t cb = ((size_t)n * e->geo.S + 15) & ~(size_t)15, mb = (size_t)n * AP_META_INTS * 4;
  AP_TRY(ap_stage(e, cb + mb, 0));
  int8_t* d_cells = (int8_t*)e->d_stage;
  int32_t* d_meta = (int32_t*)((char*)e->d_stage + cb);
  launch_boards_export(e, rows, meta, d_ids, n, d_cells, d_meta);
  AP_LAUNCH_CHECK(e);
  AP_CUDA(e, cudaMemcpyAsync(out_cells, d_cells, (size_t)n * e->geo.S, cudaMemcpyDeviceToHost, e->stream));
  return d2h_sync(e, out_meta, d_meta, mb);
}

int ap_boards_export(ap_engine* e, const int32_t* game_ids, int32_t n, int8_t* out_cells, int32_t* out_meta) {
  AP_ENTER(e);
  AP_TRY(ap_ids(e, game_ids, n));
  return export_common(e, e->rows, e->meta, e->d_ids, n, out_cells, out_meta);
}

int ap_boards_import(ap_engine* e, const int32_t* game_ids, int32_t n, const int8_t* cells, const int32_t* meta) {
  AP_ENTER(e);
  AP_TRY(ap_ids(e, game_ids, n));
  size_t cb = ((size_t)n * e->geo.S + 15) & ~(size_t)15, mb = (size_t)n * AP_META_INTS * 4;
  AP_TRY(ap_stage(e, cb + mb, 0));
  int8_t* d_cells = (int8_t*)e->d_stage;
  int32_t* d_meta = (int32_t*)((char*)e->d_stage + cb);
  AP_TRY(h2d(e, d_cells, cells, (size_t)n * e->geo.S));
  AP_TRY(h2d(e, d_meta, meta, mb));
  launch_boards_import(e, n, d_cells, d_meta);
  AP_LAUNCH_CHECK(e);
  return ap_sync(e);
}

// ---- search -------------------------------------------------------------------------------

// trees left behind by ap_pure_run are never continued (see ap_engine::pure_tree)
static int drop_pure_trees(ap_engine* e) {
  if (e->pure_tree) {
    launch_tree_reset_all(e);
    AP_LAUNCH_CHECK(e);
    e->pure_tree = false;
  }
  return AP_OK;
}

int ap_search_select(ap_engine* e, uint8_t* out_terminal, int32_t* out_depth, int16_t* out_path) {
  AP_ENTER(e);
  AP_TRY(drop_pure_trees(e));
  AP_TRY(ensure_pool(e, e->geo.S));
  launch_select(e);
  AP_LAUNCH_CHECK(e);
  const int G = e->geo.G;
  if (out_terminal) AP_CUDA(e, cudaMemcpyAsync(out_terminal, e->leaves.terminal, G, cudaMemcpyDeviceToHost, e->stream));
  if (out_depth) AP_CUDA(e, cudaMemcpyAsync(out_depth, e->leaves.depth, 4 * (size_t)G, cudaMemcpyDeviceToHost, e->stream));
  if (out_path)
    AP_CUDA(e, cudaMemcpyAsync(out_path, e->leaves.path, 2 * (size_t)G * e->geo.S, cudaMemcpyDeviceToHost, e->stream));
  return ap_sync(e);
}

int ap_search_leaf_export(ap_engine* e, int8_t* out_cells, int32_t* out_meta) {
  AP_ENTER(e);
  return export_common(e, e->leaves.rows, e->leaves.meta, nullptr, e->geo.G, out_cells, out_meta);
}

int ap_search_leaf_features(ap_engine* e, float* out) {
  AP_ENTER(e);
  size_t bytes = (size_t)e->geo.G * 9 * e->geo.S * 4;
  AP_TRY(ap_stage(e, bytes, 0));
  launch_boards_features(e, e->leaves.rows, e->leaves.meta, nullptr, e->geo.G, (float*)e->d_stage);
  e->launches++;
  AP_LAUNCH_CHECK(e);
  return d2h_sync(e, out, e->d_stage, bytes);
}

int ap_search_expand_backup(ap_engine* e, const int32_t* counts, const int16_t* acts, const double* priors,
                            const double* values) {
  AP_ENTER(e);
  if (!counts || !acts || !priors || !values) return ap_fail(e, AP_ERR_BAD_ARG, "null argument");
  const size_t G = e->geo.G, S = e->geo.S;
  for (size_t g = 0; g < G; ++g)
    if (counts[g] < 0 || counts[g] > (int)S) return ap_fail(e, AP_ERR_BAD_ARG, "counts out of range");
  size_t o_counts = 0, o_acts = o_counts + ((G * 4 + 15) & ~15ull), o_pri = o_acts + ((G * S * 2 + 15) & ~15ull),
         o_val = o_pri + G * S * 8, tot = o_val + G * 8;
  AP_TRY(ap_stage(e, tot, 0));
  char* d = (char*)e->d_stage;
  AP_TRY(h2d(e, d + o_counts, counts, G * 4));
  AP_TRY(h2d(e, d + o_acts, acts, G * S * 2));
  AP_TRY(h2d(e, d + o_pri, priors, G * S * 8));
  AP_TRY(h2d(e, d + o_val, values, G * 8));
  launch_expand_backup(e, (int32_t*)(d + o_counts), (int16_t*)(d + o_acts), (double*)(d + o_pri), (double*)(d + o_val),
                       nullptr, nullptr);
  AP_LAUNCH_CHECK(e);
  return check_errflags(e);
}

int ap_search_expand_backup_dense(ap_engine* e, const float* priors, const float* values) {
  AP_ENTER(e);
  if (!priors || !values) return ap_fail(e, AP_ERR_BAD_ARG, "null argument");
  const size_t G = e->geo.G, S = e->geo.S;
  AP_TRY(h2d(e, e->d_probs, priors, G * S * 4));
  AP_TRY(h2d(e, e->d_values, values, G * 4));
  launch_expand_backup(e, nullptr, nullptr, nullptr, nullptr, e->d_probs, e->d_values);
  AP_LAUNCH_CHECK(e);
  return check_errflags(e);
}

int ap_search_run(ap_engine* e, int32_t n_playout) {
  AP_ENTER(e);
  if (!e->net) return ap_fail(e, AP_ERR_NO_NET, "ap_search_run: no net loaded (ap_net_load)");
  AP_TRY(drop_pure_trees(e));
  AP_TRY(ensure_pool(e, (long long)n_playout * e->geo.S));
  // phases per lock-step: select, features, one per trunk conv, heads, expand/backup
  const int phases = net_phase_count(e) + 2;
  if (e->profile) {
    size_t need = (size_t)n_playout * phases + 1;
    while (e->prof_events.size() < need) {
      cudaEvent_t ev;
      AP_CUDA(e, cudaEventCreate(&ev));
      e->prof_events.push_back(ev);
    }
    e->prof_cursor = 0;
  }
  // AP_COMPACT_KERNEL=1: the unfused lock-step (A/B): separate order-preserving compaction kernel, features kernel
  // and FC finish kernel instead of the ticket + feature emission inside k_select and the softmax inside k_expand_backup
  static const bool compact_kernel = getenv("AP_COMPACT_KERNEL") && atoi(getenv("AP_COMPACT_KERNEL")) != 0;
  // AP_GRAPH_MAX_GAMES: largest batch that replays the lock-steps as a CUDA graph (default 256; 0 disables)
  static const int graph_max_games = getenv("AP_GRAPH_MAX_GAMES") ? atoi(getenv("AP_GRAPH_MAX_GAMES")) : 256;
  const bool compact = net_can_compact(e);
  // AP_ENQUEUE_CHUNK: lock-steps per chunk of the bounded launch queue (default 32; 0 = enqueue everything at once)
  static const int enqueue_chunk = getenv("AP_ENQUEUE_CHUNK") ? atoi(getenv("AP_ENQUEUE_CHUNK")) : 32;
  bool capturing = false;
  auto enqueue = [&]() -> int {
    AP_CUDA(e, cudaMemsetAsync(e->leaves.n_eval, 0, 4, e->stream));
    for (int it = 0; it < n_playout; ++it) {
      if (!capturing && enqueue_chunk > 0 && it > 0 && it % enqueue_chunk == 0) {
        const int c = it / enqueue_chunk;  // chunk about to be enqueued; chunk c - 1 was just completed on the host side
        if (!e->chunk_ev[0])
          for (auto& ev : e->chunk_ev) AP_CUDA(e, cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
        AP_CUDA(e, cudaEventRecord(e->chunk_ev[(c - 1) % 3], e->stream));
        if (c >= 2) AP_CUDA(e, cudaEventSynchronize(e->chunk_ev[(c - 2) % 3]));
      }
      launch_select(e, compact && !compact_kernel);
      AP_LAUNCH_CHECK(e);
      prof_mark(e);
      // only the non-terminal leaves are evaluated (the reference discards the evaluator's answer at a terminal
      // leaf, mcts_alphaZero.py:124-136): the net runs on the compacted batch, expand/backup reads through the slot map
      AP_TRY(net_forward_leaves(e, 0, compact, !compact_kernel));
      launch_expand_backup(e, nullptr, nullptr, nullptr, nullptr, e->d_probs, e->d_values,
                           compact ? e->leaves.slot : nullptr, compact && !compact_kernel);
      AP_LAUNCH_CHECK(e);
      prof_mark(e);
    }
    return AP_OK;
  };
  const bool use_graph = !e->profile && e->geo.G <= graph_max_games && n_playout >= 4;
  if (use_graph && (!e->run_graph || e->run_graph_playouts != n_playout || e->run_graph_gen != e->net_generation + (e->pool_generation << 32))) {
    if (e->run_graph) cudaGraphExecDestroy(e->run_graph);
    e->run_graph = nullptr;
    const uint64_t l0 = e->launches;
    cudaGraph_t g = nullptr;
    AP_CUDA(e, cudaStreamBeginCapture(e->stream, cudaStreamCaptureModeThreadLocal));
    capturing = true;
    const int rc = enqueue();
    capturing = false;
    const cudaError_t ce = cudaStreamEndCapture(e->stream, &g);
    if (rc != AP_OK) {
      if (g) cudaGraphDestroy(g);
      return rc;
    }
    AP_CUDA(e, ce);
    const cudaError_t ci = cudaGraphInstantiate(&e->run_graph, g, 0);
    cudaGraphDestroy(g);
    AP_CUDA(e, ci);
    e->run_graph_launches = e->launches - l0;
    e->launches = l0;  // counted per replay below
    e->run_graph_playouts = n_playout;
    e->run_graph_gen = e->net_generation + (e->pool_generation << 32);
  }
  AP_CUDA(e, cudaEventRecord(e->ev0, e->stream));
  prof_mark(e);
  if (use_graph) {
    AP_CUDA(e, cudaGraphLaunch(e->run_graph, e->stream));
    e->launches += e->run_graph_launches;
  } else {
    AP_TRY(enqueue());
  }
  AP_CUDA(e, cudaEventRecord(e->ev1, e->stream));
  AP_CUDA(e, cudaStreamSynchronize(e->stream));
  cudaEventElapsedTime(&e->last_total_ms, e->ev0, e->ev1);
  if (e->profile) {
    e->prof_ms.assign(phases, 0.f);
    for (int it = 0; it < n_playout; ++it)
      for (int ph = 0; ph < phases; ++ph) {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e->prof_events[(size_t)it * phases + ph], e->prof_events[(size_t)it * phases + ph + 1]);
        e->prof_ms[ph] += ms;
      }
  }
  AP_TRY(net_check_err(e));
  return check_errflags(e);
}

// ---- opt-in multi-leaf search with virtual loss (kernels and semantics: tree.cu, k_select_vl) -------------------
static int vl_prepare(ap_engine* e, int k) {
  const size_t G = e->geo.G, S = e->geo.S;
  if (e->vl_kstride < k) {
    // leaf records for G * kstride leaves (the old ones, if any, stay in e->allocs until the handle dies)
    const int ks = std::max(k, std::min(64, net_board_capacity(e) / (int)G));
    const size_t L = G * (size_t)ks;
    Leaves& lv = e->leaves_vl;
    if (e->vl_graph) {
      cudaGraphExecDestroy(e->vl_graph);
      e->vl_graph = nullptr;
    }
#define VALLOC(ptr, bytes) AP_TRY(dev_alloc(e, (void**)&(ptr), (bytes)))
    VALLOC(lv.rows, L * AP_ROWS * 4);
    VALLOC(lv.meta, L * sizeof(BoardMeta));
    VALLOC(lv.node, L * 4);
    VALLOC(lv.terminal, L);
    VALLOC(lv.winner, L);
    VALLOC(lv.depth, L * 4);
    VALLOC(lv.path, L * S * 2);
    VALLOC(lv.slot, L * 4);
    VALLOC(lv.game_of_slot, L * 4);
    VALLOC(lv.n_eval, 4);
    if (!e->vl_remain) {
      VALLOC(e->vl_remain, G * 4);
      VALLOC(e->vl_issued, G * 4);
    }
#undef VALLOC
    lv.active = e->leaves.active;
    e->vl_kstride = ks;
  }
  if (!e->vn || e->vn_cap != e->geo.cap) {  // follows the pools when they grow; all zero between lock-steps
    if (e->vn) dev_free(e, e->vn, G * (size_t)e->vn_cap * 4);
    e->vn = nullptr;
    AP_TRY(dev_alloc(e, (void**)&e->vn, G * (size_t)e->geo.cap * 4));
    e->vn_cap = e->geo.cap;
  }
  return AP_OK;
}

int ap_search_run_vl(ap_engine* e, int32_t n_playout, int32_t k) {
  AP_ENTER(e);
  if (!e->net) return ap_fail(e, AP_ERR_NO_NET, "ap_search_run_vl: no net loaded (ap_net_load)");
  if (n_playout < 1 || k < 1) return ap_fail(e, AP_ERR_BAD_ARG, "ap_search_run_vl: n_playout >= 1, k >= 1");
  if (!net_can_compact(e)) return ap_fail(e, AP_ERR_BAD_ARG, "ap_search_run_vl: this net has no compacted-batch path");
  if ((long long)k * e->geo.G > net_board_capacity(e) || k > 64)
    return ap_fail(e, AP_ERR_BAD_ARG, "ap_search_run_vl: n_games * k exceeds the net batch (" +
                                          std::to_string(net_board_capacity(e)) + " boards) or k > 64");
  AP_TRY(drop_pure_trees(e));
  AP_TRY(ensure_pool(e, (long long)n_playout * e->geo.S));
  AP_TRY(vl_prepare(e, k));
  const int G = e->geo.G;
  std::vector<int32_t> rem(G, n_playout);
  AP_TRY(h2d(e, e->vl_remain, rem.data(), sizeof(int32_t) * G));
  AP_CUDA(e, cudaStreamSynchronize(e->stream));
  // a fresh root takes one lock-step for its first playout, then ceil((n - 1) / k); a reused tree ceil(n / k) <= that
  const int steps = 1 + (n_playout - 1 + k - 1) / k;
  auto enqueue = [&]() -> int {
    AP_CUDA(e, cudaMemsetAsync(e->leaves_vl.n_eval, 0, 4, e->stream));
    for (int it = 0; it < steps; ++it) {
      launch_select_vl(e, e->leaves_vl, e->vn, k, e->vl_kstride, e->vl_remain, e->vl_issued);
      AP_LAUNCH_CHECK(e);
      AP_TRY(net_run_compacted(e, G * k, e->leaves_vl.n_eval));
      launch_expand_backup_vl(e, e->leaves_vl, e->vn, e->vl_kstride, e->vl_issued);
      AP_LAUNCH_CHECK(e);
    }
    return AP_OK;
  };
  static const int graph_max_games = getenv("AP_GRAPH_MAX_GAMES") ? atoi(getenv("AP_GRAPH_MAX_GAMES")) : 256;
  const bool use_graph = !e->profile && G <= graph_max_games && steps >= 4;
  const uint64_t gen = e->net_generation + (e->pool_generation << 32);
  if (use_graph && (!e->vl_graph || e->vl_graph_playouts != n_playout || e->vl_graph_k != k || e->vl_graph_gen != gen)) {
    if (e->vl_graph) cudaGraphExecDestroy(e->vl_graph);
    e->vl_graph = nullptr;
    const uint64_t l0 = e->launches;
    cudaGraph_t g = nullptr;
    AP_CUDA(e, cudaStreamBeginCapture(e->stream, cudaStreamCaptureModeThreadLocal));
    const int rc = enqueue();
    const cudaError_t ce = cudaStreamEndCapture(e->stream, &g);
    if (rc != AP_OK) {
      if (g) cudaGraphDestroy(g);
      return rc;
    }
    AP_CUDA(e, ce);
    const cudaError_t ci = cudaGraphInstantiate(&e->vl_graph, g, 0);
    cudaGraphDestroy(g);
    AP_CUDA(e, ci);
    e->vl_graph_launches = e->launches - l0;
    e->launches = l0;
    e->vl_graph_playouts = n_playout;
    e->vl_graph_k = k;
    e->vl_graph_gen = gen;
  }
  AP_CUDA(e, cudaEventRecord(e->ev0, e->stream));
  if (use_graph) {
    AP_CUDA(e, cudaGraphLaunch(e->vl_graph, e->stream));
    e->launches += e->vl_graph_launches;
  } else {
    AP_TRY(enqueue());
  }
  AP_CUDA(e, cudaEventRecord(e->ev1, e->stream));
  AP_CUDA(e, cudaStreamSynchronize(e->stream));
  cudaEventElapsedTime(&e->last_total_ms, e->ev0, e->ev1);
  AP_TRY(net_check_err(e));
  return check_errflags(e);
}

int ap_search_profile(ap_engine* e, int32_t enable, float* out_ms, int32_t cap) {
  AP_ENTER(e);
  e->profile = enable;
  int n = (int)e->prof_ms.size();
  if (out_ms)
    for (int i = 0; i < n && i < cap; ++i) out_ms[i] = e->prof_ms[i];
  return n;
}

int ap_search_timing(ap_engine* e, float* out_total_ms, float* out_net_ms) {
  AP_ENTER(e);
  if (out_total_ms) *out_total_ms = e->last_total_ms;
  if (out_net_ms) *out_net_ms = e->last_net_ms;
  return AP_OK;
}

int ap_search_root(ap_engine* e, const int32_t* game_ids, int32_t n, int32_t* out_count, int16_t* out_acts,
                   int32_t* out_visits, double* out_q, int32_t* out_root_n) {
  AP_ENTER(e);
  AP_TRY(ap_ids(e, game_ids, n));
  const size_t S = e->geo.S, nn = n;
  size_t o_cnt = 0, o_rn = o_cnt + ((nn * 4 + 15) & ~15ull), o_acts = o_rn + ((nn * 4 + 15) & ~15ull),
         o_vis = o_acts + ((nn * S * 2 + 15) & ~15ull), o_q = o_vis + ((nn * S * 4 + 15) & ~15ull), tot = o_q + nn * S * 8;
  AP_TRY(ap_stage(e, tot, 0));
  char* d = (char*)e->d_stage;
  launch_root(e, e->d_ids, n, (int32_t*)(d + o_cnt), (int16_t*)(d + o_acts), (int32_t*)(d + o_vis),
              out_q ? (double*)(d + o_q) : nullptr, (int32_t*)(d + o_rn));
  AP_LAUNCH_CHECK(e);
  AP_CUDA(e, cudaMemcpyAsync(out_count, d + o_cnt, nn * 4, cudaMemcpyDeviceToHost, e->stream));
  AP_CUDA(e, cudaMemcpyAsync(out_acts, d + o_acts, nn * S * 2, cudaMemcpyDeviceToHost, e->stream));
  AP_CUDA(e, cudaMemcpyAsync(out_visits, d + o_vis, nn * S * 4, cudaMemcpyDeviceToHost, e->stream));
  if (out_q) AP_CUDA(e, cudaMemcpyAsync(out_q, d + o_q, nn * S * 8, cudaMemcpyDeviceToHost, e->stream));
  if (out_root_n) AP_CUDA(e, cudaMemcpyAsync(out_root_n, d + o_rn, nn * 4, cudaMemcpyDeviceToHost, e->stream));
  return ap_sync(e);
}

int ap_search_root_probs(ap_engine* e, double temp, double* out) {
  AP_ENTER(e);
  if (!(temp > 0)) return ap_fail(e, AP_ERR_BAD_ARG, "temp must be > 0");
  size_t bytes = (size_t)e->geo.G * e->geo.S * 8;
  AP_TRY(ap_stage(e, bytes, 0));
  launch_root_probs(e, temp, (double*)e->d_stage);
  AP_LAUNCH_CHECK(e);
  return d2h_sync(e, out, e->d_stage, bytes);
}

int ap_selfplay_pick(ap_engine* e, double temp, double eps, double alpha, uint64_t seed, uint32_t ply, int32_t* out_moves,
                     float* out_pi, double* out_noise) {
  AP_ENTER(e);
  if (!(temp > 0) || eps < 0 || eps > 1 || !(alpha > 0)) return ap_fail(e, AP_ERR_BAD_ARG, "temp > 0, 0 <= eps <= 1, alpha > 0");
  if (!out_moves || !out_pi) return ap_fail(e, AP_ERR_BAD_ARG, "null argument");
  const size_t G = e->geo.G, S = e->geo.S;
  const size_t mb = (4 * G + 15) & ~15ull, pb = 4 * G * S, nb = out_noise ? 8 * G * S : 0;
  AP_TRY(ap_stage(e, mb + pb + nb, 0));
  int32_t* d_m = (int32_t*)e->d_stage;
  float* d_p = (float*)((char*)e->d_stage + mb);
  double* d_n = out_noise ? (double*)((char*)e->d_stage + mb + pb) : nullptr;
  launch_selfplay_pick(e, temp, eps, alpha, seed, ply, d_m, d_p, d_n);
  AP_LAUNCH_CHECK(e);
  AP_TRY(traj_append_pick(e, d_p));  // device-side trajectories (ap_traj_create): this ply's (state, pi, player)
  AP_CUDA(e, cudaMemcpyAsync(out_moves, d_m, 4 * G, cudaMemcpyDeviceToHost, e->stream));
  if (out_noise) AP_CUDA(e, cudaMemcpyAsync(out_noise, d_n, nb, cudaMemcpyDeviceToHost, e->stream));
  return d2h_sync(e, out_pi, d_p, pb);
}

int ap_search_advance(ap_engine* e, const int32_t* game_ids, int32_t n, const int32_t* moves) {
  AP_ENTER(e);
  AP_TRY(drop_pure_trees(e));
  AP_TRY(ap_ids(e, game_ids, n));
  AP_TRY(ap_stage(e, sizeof(int32_t) * n, 0));
  AP_TRY(h2d(e, e->d_stage, moves, sizeof(int32_t) * n));
  launch_advance(e, n, (int32_t*)e->d_stage);
  AP_LAUNCH_CHECK(e);
  return ap_sync(e);
}

int ap_search_set_active(ap_engine* e, const uint8_t* active) {
  AP_ENTER(e);
  uint8_t* d = const_cast<uint8_t*>(e->leaves.active);
  if (!active) {
    AP_CUDA(e, cudaMemsetAsync(d, 1, e->geo.G, e->stream));
  } else {
    std::vector<uint8_t> a(e->geo.G);
    for (int g = 0; g < e->geo.G; ++g) a[g] = active[g] ? 1 : 0;
    AP_CUDA(e, cudaMemcpyAsync(d, a.data(), a.size(), cudaMemcpyHostToDevice, e->stream));
  }
  return ap_sync(e);
}

int ap_search_stats(ap_engine* e, uint64_t* out5) {
  AP_ENTER(e);
  unsigned long long h[8];
  AP_TRY(d2h_sync(e, h, e->stats, sizeof(h)));
  for (int i = 0; i < 6; ++i) out5[i] = h[i];
  AP_CUDA(e, cudaMemsetAsync(e->stats, 0, sizeof(h), e->stream));
  return ap_sync(e);
}

// ---- mcts_pure ------------------------------------------------------------------------------

int ap_pure_run(ap_engine* e, int32_t n_playout, uint64_t seed, int32_t rollout_mode, int32_t* out_move) {
  AP_ENTER(e);
  if (rollout_mode < 0 || rollout_mode > 2) return ap_fail(e, AP_ERR_BAD_ARG, "rollout_mode must be 0, 1 or 2");
  AP_TRY(ap_stage(e, 4 * (size_t)e->geo.G, 0));
  if (e->cap_auto && 1ll + (long long)n_playout * e->geo.S > e->geo.cap) {
    // the fused kernel starts every game from a fresh root: drop the trees, then grow the empty pools
    launch_tree_reset_all(e);
    AP_LAUNCH_CHECK(e);
    AP_TRY(ensure_pool(e, (long long)n_playout * e->geo.S));
  }
  AP_CUDA(e, cudaEventRecord(e->ev0, e->stream));
  launch_pure_run(e, n_playout, seed, rollout_mode, (int32_t*)e->d_stage);
  AP_LAUNCH_CHECK(e);
  e->pure_tree = true;
  AP_CUDA(e, cudaEventRecord(e->ev1, e->stream));
  AP_TRY(d2h_sync(e, out_move, e->d_stage, 4 * (size_t)e->geo.G));
  cudaEventElapsedTime(&e->last_total_ms, e->ev0, e->ev1);
  return check_errflags(e);
}

int ap_rollout_eval(ap_engine* e, uint64_t seed, int8_t* out_value, int16_t* out_plies) {
  return ap_rollout_eval2(e, seed, 0, out_value, out_plies);
}

int ap_rollout_eval2(ap_engine* e, uint64_t seed, int32_t impl, int8_t* out_value, int16_t* out_plies) {
  AP_ENTER(e);
  if (impl != 0 && impl != 2) return ap_fail(e, AP_ERR_BAD_ARG, "rollout impl must be 0 (permutation) or 2 (ply by ply)");
  const size_t G = e->geo.G;
  AP_TRY(ap_stage(e, 4 * G, 0));
  int8_t* d_v = (int8_t*)e->d_stage;
  int16_t* d_p = (int16_t*)((char*)e->d_stage + ((G + 15) & ~15ull));
  AP_TRY(ap_stage(e, ((G + 15) & ~15ull) + 2 * G, 0));
  d_v = (int8_t*)e->d_stage;
  d_p = (int16_t*)((char*)e->d_stage + ((G + 15) & ~15ull));
  launch_rollout_eval(e, seed, impl, nullptr, d_v, d_p);
  AP_LAUNCH_CHECK(e);
  AP_CUDA(e, cudaMemcpyAsync(out_value, d_v, G, cudaMemcpyDeviceToHost, e->stream));
  return d2h_sync(e, out_plies, d_p, 2 * G);
}

int ap_rollout_eval_keys(ap_engine* e, const uint32_t* keys, int8_t* out_value, int16_t* out_plies) {
  AP_ENTER(e);
  if (!keys || !out_value || !out_plies) return ap_fail(e, AP_ERR_BAD_ARG, "null argument");
  if (e->geo.W > 15) return ap_fail(e, AP_ERR_BAD_ARG, "the permutation rollout needs width <= 15");
  const size_t G = e->geo.G, kb = G * 256 * sizeof(uint32_t), vb = (G + 15) & ~15ull;
  AP_TRY(ap_stage(e, kb + vb + 2 * G, 0));
  uint32_t* d_k = (uint32_t*)e->d_stage;
  int8_t* d_v = (int8_t*)e->d_stage + kb;
  int16_t* d_p = (int16_t*)((char*)e->d_stage + kb + vb);
  AP_TRY(h2d(e, d_k, keys, kb));
  launch_rollout_eval(e, 0, 0, d_k, d_v, d_p);
  AP_LAUNCH_CHECK(e);
  AP_CUDA(e, cudaMemcpyAsync(out_value, d_v, G, cudaMemcpyDeviceToHost, e->stream));
  return d2h_sync(e, out_plies, d_p, 2 * G);
}

int ap_rollout_hash(ap_engine* e, int8_t* out_value) {
  AP_ENTER(e);
  AP_TRY(ap_stage(e, e->geo.G, 0));
  launch_rollout_hash(e, (int8_t*)e->d_stage);
  AP_LAUNCH_CHECK(e);
  return d2h_sync(e, out_value, e->d_stage, e->geo.G);
}

}  // extern "C"
