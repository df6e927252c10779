/*
 * alphapig_b200.h -- C ABI of libalphapig_b200.so (sm_100a only, no CPU fallback).
 *
 * The reference (anxingle/AlphaPig) has no FFI: its "operator API" for the
 * self-play hot path is a set of duck-typed Python call signatures.  Each entry
 * point below names the reference interface it replaces (file:line relative to
 * the reference tree).  The Python shims in alphapig_b200/ bind these with
 * ctypes and re-expose the reference's own class/method names; INTEGRATION.md
 * shows the binding a reference maintainer would add.
 *
 * Conventions
 *  - every function returns int: AP_OK (0) or a negative ap_status; the text of
 *    the last failure on a handle is ap_last_error(e).
 *  - the library owns all device memory; host pointers are borrowed for the
 *    duration of the call; calls are synchronous on return unless the name ends
 *    in _async (then ap_sync()).
 *  - not re-entrant per handle: one handle per GPU, one host thread per handle.
 *  - game_ids == NULL means games 0..n-1.
 *  - a move is h*width + w (reference game.py:46-56); players are 1 and 2.
 *  - S = width*height <= 256, width,height <= 16.
 */
#ifndef ALPHAPIG_B200_H
#define ALPHAPIG_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ap_engine ap_engine;

typedef enum {
  AP_OK = 0,
  AP_ERR_BAD_ARG = -1,        /* reference: Exception (game.py:36-38, 206-208)          */
  AP_ERR_ILLEGAL_MOVE = -2,   /* reference: ValueError from list.remove (game.py:120)   */
  AP_ERR_POOL_EXHAUSTED = -3, /* node pool of a game is full (no reference analogue)    */
  AP_ERR_CUDA = -4,
  AP_ERR_NO_NET = -5,         /* ap_search_run / ap_net_forward before ap_net_load      */
  AP_ERR_BAD_HANDLE = -6
} ap_status;

typedef struct {
  int32_t width, height, n_in_row; /* Board(width,height,n_in_row)      game.py:24-33   */
  int32_t n_games;                 /* G concurrent games on this GPU                     */
  int32_t node_capacity;           /* tree nodes per game (0 = pick from n_playout_hint) */
  int32_t n_playout_hint;          /* MCTS(n_playout)                mcts_alphaZero.py:93 */
  int32_t device;                  /* CUDA device ordinal                                */
  int32_t flags;                   /* AP_FLAG_*                                          */
  double  c_puct;                  /* MCTS(c_puct)                   mcts_alphaZero.py:93 */
} ap_config;

#define AP_FLAG_NONE 0
/* the handle's CUDA stream gets the highest priority: for small latency-sensitive handles (a trainer's batch forward and
 * replay ring) that share the GPU with another handle's search - their kernels are scheduled as soon as an SM frees up
 * instead of queueing behind the search's next persistent kernel */
#define AP_FLAG_HIGH_PRIORITY_STREAM 1
#define AP_META_INTS 8 /* current_player, last_move, n_stones, hist0..3 (most recent first, -1 = none), start_player */

/* ---- lifecycle ------------------------------------------------------------ */
int ap_engine_create(const ap_config* cfg, ap_engine** out);
int ap_engine_destroy(ap_engine* e);
const char* ap_last_error(const ap_engine* e);
const char* ap_version(void);
int ap_sync(ap_engine* e);
/* bytes of device memory held by the handle */
int ap_engine_memory(const ap_engine* e, uint64_t* out_bytes);

/* tree nodes per game the pools hold now.  cfg.node_capacity > 0 is a hard limit (AP_ERR_POOL_EXHAUSTED when a game
 * runs over it); cfg.node_capacity <= 0 lets the library pick 2 * n_playout_hint * S + S + 2 and GROW the pools
 * whenever a search could run out (a re-rooted subtree keeps its visits, mcts_alphaZero.py:159-167, so no fixed size is
 * safe for self-play with tree reuse; the reference's trees are unbounded Python objects). */
int ap_engine_node_capacity(const ap_engine* e, int32_t* out_nodes);

/* ---- boards: replaces game.Board (game.py:21-170) --------------------------- */
/* Board.init_board(start_player)                                   game.py:35-44 */
int ap_boards_reset(ap_engine* e, const int32_t* game_ids, int32_t n, const int32_t* start_player);
/* Board.do_move(move); out_status[i] = AP_OK / AP_ERR_ILLEGAL_MOVE (board untouched)  game.py:117-125 */
int ap_boards_do_move(ap_engine* e, const int32_t* game_ids, const int32_t* moves, int32_t n, int32_t* out_status);
/* Board.game_end() -> (end, winner in {1,2,-1})                  game.py:127-167 */
int ap_boards_status(ap_engine* e, const int32_t* game_ids, int32_t n, uint8_t* out_end, int8_t* out_winner);
/* Board.availables as a bit mask, bit m of out_mask[i][m/32]          game.py:41 */
int ap_boards_legal(ap_engine* e, const int32_t* game_ids, int32_t n, uint32_t* out_mask /* [n][8] */);
/* Board.current_state() -> float32 [n][9][width][height], flip included  game.py:68-94 */
int ap_boards_features(ap_engine* e, const int32_t* game_ids, int32_t n, float* out);
/* the same planes bit-packed, np.packbits order: uint8 [n][ceil(9*S/8)] (the record format of ap_replay_push) */
int ap_boards_features_packed(ap_engine* e, const int32_t* game_ids, int32_t n, uint8_t* out);
/* Board.states / history / current_player / last_move                game.py:24-44 */
int ap_boards_export(ap_engine* e, const int32_t* game_ids, int32_t n, int8_t* out_cells /* [n][S] 0,1,2 */,
                     int32_t* out_meta /* [n][AP_META_INTS] */);
int ap_boards_import(ap_engine* e, const int32_t* game_ids, int32_t n, const int8_t* cells, const int32_t* meta);

/* ---- search: replaces mcts_alphaZero.MCTS / TreeNode (mcts_alphaZero.py:19-170) -- */
/* One lock-step of MCTS._playout up to the evaluator call (:115-124): every game
 * descends from its root by TreeNode.select (fp64 PUCT, first max) applying
 * Board.do_move to a scratch copy, and stops at a leaf.
 *   out_terminal[g] : 1 if game_end() at the leaf (evaluator result is discarded, :126-136)
 *   out_depth[g]    : plies descended; out_path[g][0..depth) the moves taken        */
int ap_search_select(ap_engine* e, uint8_t* out_terminal /* [G] or NULL */, int32_t* out_depth /* [G] or NULL */,
                     int16_t* out_path /* [G][S] or NULL */);
/* Leaf boards of the last ap_search_select, same layout as ap_boards_export / _features. */
int ap_search_leaf_export(ap_engine* e, int8_t* out_cells, int32_t* out_meta);
int ap_search_leaf_features(ap_engine* e, float* out /* [G][9][W][H] */);
/* Second half of _playout (:126-139): TreeNode.expand with the evaluator's
 * (action, prior) list in the given order (counts[g] entries of acts/priors row g),
 * terminal override of the value, then update_recursive(-leaf_value).           */
int ap_search_expand_backup(ap_engine* e, const int32_t* counts /* [G] */, const int16_t* acts /* [G][S] */,
                            const double* priors /* [G][S] */, const double* values /* [G] */);
/* Same, priors given densely by move index; children = legal moves ascending
 * (what PolicyValueNet.policy_value_fn returns, policy_value_net_mxnet_simple.py:207-226). */
int ap_search_expand_backup_dense(ap_engine* e, const float* priors /* [G][S] */, const float* values /* [G] */);
/* MCTS.get_move_probs (:141-157) with the device net as policy_value_fn:
 * n_playout x (select -> features -> net -> expand/backup), no host round trip. */
int ap_search_run(ap_engine* e, int32_t n_playout);
/* OPT-IN multi-leaf search for small batches (interactive play: human_play_mxnet.py, evaluate/ChessClient.py run ONE
 * game, which is launch-latency bound in the strictly sequential form above): up to k playouts of every game are in
 * flight per lock-step, kept apart by virtual loss (every in-flight visit of a child counts as a loss until its
 * backup), and evaluated as one net batch.  n_games * k must fit the net batch (256 boards for small engines).
 * k = 1 builds the same tree as ap_search_run bit for bit; k > 1 changes visit counts (NOT the reference's search:
 * mcts_alphaZero.py:147-149 runs the playouts one after the other) while every game still gets exactly n_playout
 * playouts. */
int ap_search_run_vl(ap_engine* e, int32_t n_playout, int32_t k);
/* Root children in insertion order: acts, visit counts, Q; root's own N.   (:152-154) */
int ap_search_root(ap_engine* e, const int32_t* game_ids, int32_t n, int32_t* out_count, int16_t* out_acts /* [n][S] */,
                   int32_t* out_visits /* [n][S] */, double* out_q /* [n][S] or NULL */, int32_t* out_root_n /* [n] or NULL */);
/* softmax(1/temp * log(visits + 1e-10)) scattered by move index, fp64        (:13-16,155) */
int ap_search_root_probs(ap_engine* e, double temp, double* out /* [G][S] */);
/* MCTSPlayer.get_action(board, temp, return_prob=1) with is_selfplay=1 for EVERY game, sampled on the device
 * (:187-215): move ~ (1-eps) * pi + eps * Dirichlet(alpha) with pi = softmax(1/temp * log(visits + 1e-10));
 * out_pi [G][S] fp32 is the un-noised pi scattered by move index (the training record); out_noise (nullable, tests)
 * the Dirichlet sample.  Randomness: Philox streams keyed by (seed, game, ply); the reference uses eps 0.25,
 * alpha 0.3.  A full board gives move -1.  Does not advance the tree (ap_search_advance does). */
int ap_selfplay_pick(ap_engine* e, double temp, double eps, double alpha, uint64_t seed, uint32_t ply,
                     int32_t* out_moves /* [G] */, float* out_pi /* [G][S] */, double* out_noise /* [G][S] or NULL */);
/* MCTS.update_with_move(move): re-root on the child (subtree kept) or fresh root (-1 / absent)  (:159-167) */
int ap_search_advance(ap_engine* e, const int32_t* game_ids, int32_t n, const int32_t* moves);
/* Restrict every following search (ap_search_select / _run, ap_pure_run) to the games with active[g] != 0; NULL =
 * all games (the default).  Skipped games keep their boards and trees and report move -1 from ap_pure_run.  The
 * reference plays its arena games one after the other (train_mxnet.py:239-263) and simply stops calling get_action
 * for a finished game; a batch of concurrent games needs this mask for the same effect. */
int ap_search_set_active(ap_engine* e, const uint8_t* active /* [G] or NULL */);
/* counters since the last call: playouts, sum of children scanned, children written, path nodes, terminal leaves */
int ap_search_stats(ap_engine* e, uint64_t* out5);

/* ---- mcts_pure (mcts_pure.py:13-206) ---------------------------------------- */
/* MCTS.get_move: n_playout x (_playout with uniform priors + random rollout) fully on
 * device, one CTA per game; out_move[g] = first max by visits (:159-169).
 * rollout_mode 0 = uniform random legal moves (rollout_policy_fn, :13-17) drawn as one random permutation of
 *                  the empty cells per rollout + a bit-descent to the first completed line (same distribution
 *                  of result and length as playing ply by ply; see csrc/rollout.cu),
 *              1 = deterministic position hash in {-1,0,1} (bookkeeping-parity tests),
 *              2 = the same uniform random game played ply by ply (cross-check of mode 0; also used
 *                  for 16-wide boards). */
int ap_pure_run(ap_engine* e, int32_t n_playout, uint64_t seed, int32_t rollout_mode, int32_t* out_move /* [G] */);
/* _evaluate_rollout (:138-157) from every root board: winner-from-leaf-player value and plies played */
int ap_rollout_eval(ap_engine* e, uint64_t seed, int8_t* out_value /* [G] */, int16_t* out_plies /* [G] */);
/* same with the rollout implementation chosen: impl 0 = permutation (default), 2 = ply by ply */
int ap_rollout_eval2(ap_engine* e, uint64_t seed, int32_t impl, int8_t* out_value /* [G] */, int16_t* out_plies /* [G] */);
/* the permutation rollout with INJECTED randomness: keys[g][row*16 + column] (low 24 bits used) is the draw of
 * that cell; the empty cells are played in ascending (key, cell) order - exactly what the oracle replays move by
 * move (mcts_pure.py:138-157 with the arg-max draws replaced by this order).  width <= 15. */
int ap_rollout_eval_keys(ap_engine* e, const uint32_t* keys /* [G][256] */, int8_t* out_value /* [G] */,
                         int16_t* out_plies /* [G] */);
/* the deterministic hash used by rollout_mode 1, for the host-side oracle */
int ap_rollout_hash(ap_engine* e, int8_t* out_value /* [G] */);

/* ---- policy/value net: replaces PolicyValueNet forward (policy_value_net_mxnet{,_simple}.py) -- */
typedef struct {
  const char* name;   /* reference parameter name, e.g. "conv1_weight", "bnA3_moving_var" */
  const float* data;  /* host fp32, reference shape/layout (O,I,kh,kw) / (out,in)         */
  int64_t numel;
} ap_tensor;
#define AP_ARCH_SIMPLE 0 /* policy_value_net_mxnet_simple.py:68-92 */
#define AP_ARCH_RESNET 1 /* policy_value_net_mxnet.py:70-102       */
/* builder-defined board-sized Inception-ResNet variant (BASELINE configs[3]): 3x3 stem + n_blocks x block35
 * (inception-resnet-v2.py:41-58) + the reference heads; the reference file itself is an unwired ImageNet symbol.
 * Parameter names: incep_conv1_*, b35_<i>_{t0,t1a,t1b,t2a,t2b,t2c,up}_* (alphapig_b200/params.py). n_filter = 128. */
#define AP_ARCH_INCEPTION 2
/* OR into `arch` (residual net only): every activation and weight is carried as a hi + lo fp16 pair and the
 * tensor cores compute hi*hi + lo*hi + hi*lo (near-fp32 accuracy at 3x the MMA work).  The 10-block net of
 * train_mxnet.py:79-91 needs it to stay within 1e-3 of fp32; plain fp16 operands reach 1.8e-3 there. */
#define AP_NET_SPLIT 0x100
/* OR into `arch` instead of AP_NET_SPLIT (residual net only): activations as hi + lo fp16 pairs, weights as ONE fp16
 * value rounded by error diffusion along K (the rounding errors of an output channel sum to < 1 ulp): hi*w + lo*w,
 * 2x the MMA work; 4.7e-4 on the 10-block net.  The default of the Python shims for deep residual nets. */
#define AP_NET_SPLIT_ACT 0x200
/* set_params(arg_params, aux_params)                 policy_value_net_mxnet_simple.py:33-37 */
int ap_net_load(ap_engine* e, int32_t arch, int32_t n_blocks, int32_t n_filter, const ap_tensor* tensors, int32_t n_tensors);
/* PolicyValueNet.policy_value(state_batch): host fp32 states [B][9][H][W] -> probs [B][S], values [B]   (:178-188) */
int ap_net_forward(ap_engine* e, const float* states, int32_t B, float* out_probs, float* out_values);
/* same through the independent fp32 CUDA-core kernels (slow; on-device cross-check of the tensor-core path) */
int ap_net_forward_precise(ap_engine* e, const float* states, int32_t B, float* out_probs, float* out_values);
/* same on the engine's current leaf boards (device resident), results stay on device; precise=1 runs
 * the fp32 CUDA-core reference kernels instead of the fp16 tensor-core path. */
int ap_net_forward_leaves(ap_engine* e, int32_t precise, float* out_probs /* [G][S] or NULL */, float* out_values /* [G] or NULL */);
/* device pointer + element count of the flat fp32 master weights (shared with the PyTorch
 * train step and the post-train ncclBroadcast); call ap_net_refresh after writing to it. */
int ap_net_weights_ptr(ap_engine* e, void** out_dev_ptr, int64_t* out_numel);
int ap_net_refresh(ap_engine* e);
/* name/offset table of the flat buffer: returns number of tensors; fills up to cap entries */
int ap_net_layout(ap_engine* e, int32_t cap, const char** out_names, int64_t* out_offsets, int64_t* out_numels);
/* timing of the last ap_search_run in ms (whole loop, net kernels only), measured with CUDA events on the engine stream */
int ap_search_timing(ap_engine* e, float* out_total_ms, float* out_net_ms);
/* enable/disable per-phase CUDA-event timing of ap_search_run and fetch the accumulated ms of the last run;
 * phases per lock-step: select, features, one per trunk conv layer, heads, expand/backup.  Returns #phases. */
int ap_search_profile(ap_engine* e, int32_t enable, float* out_ms, int32_t cap);
/* number of kernels this library launched on the handle since creation */
int ap_launch_count(const ap_engine* e, uint64_t* out);

/* ---- replay ring: replaces TrainPipeline.data_buffer + get_equi_data + random.sample ------------------
 * (train_mxnet.py:57 deque(maxlen=buffer_size); :115-135 get_equi_data; :153,180 data_buffer.extend;
 *  :196-199 random.sample).  One packed record per position lives in HBM; the 8 rotations / flips the
 * reference materialises in its deque are applied by the gather kernel.  A logical deque index j
 * (0 = oldest) addresses augmented sample (total - len + j): record (a / 8), symmetry (a % 8) in the
 * reference's order  [rot90^1, rot90^1+fliplr, rot90^2, rot90^2+fliplr, ...].  Square boards only. */
/* deque(maxlen): maxlen counts AUGMENTED samples, as buffer_size does in the reference */
int ap_replay_create(ap_engine* e, int64_t maxlen);
/* data_buffer.extend(get_equi_data(play_data)) for n positions: state_bits = np.packbits of the (9,H,W) 0/1
 * planes of Board.current_state() [n][ceil(9S/8)], pi [n][S], z [n] (host buffers) */
int ap_replay_push(ap_engine* e, const uint8_t* state_bits, const float* pi, const float* z, int32_t n);
/* len(data_buffer) and the number of augmented samples ever appended */
int ap_replay_size(ap_engine* e, int64_t* out_len, int64_t* out_total);
/* [data_buffer[j] for j in idx] as dense fp32 batches [B][9][H][W], [B][S], [B]; out_on_device != 0: the three
 * outputs are DEVICE pointers (e.g. torch tensors feeding train_step with no host copy) */
int ap_replay_gather(ap_engine* e, const int64_t* idx, int32_t B, float* out_states, float* out_pi, float* out_z,
                     int32_t out_on_device);

/* SGF bootstrap: Game.start_self_play(player, sgf_home, file_name) (game.py:233-304) for n_games recorded games
 * at once, records written straight into the ring (8 augmented samples per ply): state = Board.current_state()
 * before the move, pi = 0.99999 at the recorded move and 1e-6 elsewhere (:249-251), z = +-1 from `winners`
 * (1, 2 or -1).  moves [n_games][max_len] (seq_num_list, utils/sgf_dataIter.py:36), lengths [n_games].
 * out_warning[g] = 1 when a recorded move is illegal: that game contributes nothing (reference: returns
 * warning=1 and no data, :262-266). */
int ap_replay_push_sgf(ap_engine* e, const int16_t* moves, int32_t max_len, const int32_t* lengths, const int8_t* winners,
                       int32_t n_games, uint8_t* out_warning /* [n_games] or NULL */);

/* the same append from packed records - [ceil(9S/8) state bytes, zero padded to a multiple of 4][S x fp32 pi][fp32 z]
 * each - on the host (on_device == 0) or on this GPU (on_device != 0: an outbox or an NCCL all-gather result) */
int ap_replay_push_packed(ap_engine* e, const void* records, int64_t n, int32_t on_device);

/* ---- self-play trajectories on the device: replaces the states / mcts_probs / current_players lists of
 * Game_AI.start_self_play (game_ai.py:75,113-131) for all G concurrent games --------------------------------------
 * After ap_traj_create every ap_selfplay_pick also appends the ply's record - Board.current_state() of the board
 * the move is picked for, the un-noised pi, the player to move - to the game's trajectory in HBM (max_plies <= 0: S
 * plies per game).  ap_traj_finish moves the records of finished games to the OUTBOX, a flat device array of packed
 * records in the format above, with z = +1 / -1 for the plies of the winner / loser and 0 for a tie (winners[i] in
 * {1, 2, -1}, :124-128), in the order of game_ids.  The outbox is what travels: ap_traj_outbox returns its device
 * pointer and record count (valid until the next ap_traj_finish), e.g. as the send buffer of an NCCL all-gather whose
 * result the trainer rank hands to ap_replay_push_packed - no host copy of any record. */
int ap_traj_create(ap_engine* e, int32_t max_plies, int64_t outbox_records);
/* the forced random two-ply opening of game_ai.py:78-111: record (state, pi = 0.99999 at moves[i] / 1e-6 elsewhere,
 * player) for game_ids[i] BEFORE the caller plays moves[i] with ap_boards_do_move */
int ap_traj_append_forced(ap_engine* e, const int32_t* game_ids, int32_t n, const int32_t* moves);
int ap_traj_finish(ap_engine* e, const int32_t* game_ids, int32_t n, const int8_t* winners);
/* forget the plies recorded so far for these games (positions loaded from outside) */
int ap_traj_discard(ap_engine* e, const int32_t* game_ids, int32_t n);
int ap_traj_outbox(ap_engine* e, void** out_dev_ptr, int64_t* out_records, int32_t* out_record_bytes);
int ap_traj_outbox_clear(ap_engine* e);

#ifdef __cplusplus
}
#endif
#endif /* ALPHAPIG_B200_H */
