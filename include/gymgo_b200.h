/*
 * gymgo_b200.h - C ABI of the B200-native batched Go engine (libgymgo_b200.so).
 *
 * This is the drop-in boundary for GymGo's hot path.  The reference has no FFI: its seam is the
 * module-level pure-function API of gym_go/gogame.py that GoEnv reaches through `GoEnv.gogame`
 * (gym_go/envs/go_env.py:21-22).  Each entry point below names the reference function(s) it replaces;
 * the Python binding a maintainer would add is gymgo_b200/_cabi.py (ctypes), shown in INTEGRATION.md.
 *
 * Conventions
 *   - Plain C types only.  Every pointer is a DEVICE pointer into memory owned by the caller (e.g.
 *     torch.Tensor.data_ptr()); `stream` is a cudaStream_t passed as void*.  Record and dense buffers
 *     must be 16-byte aligned.  No allocation, no global state, no host synchronisation inside: calls
 *     only enqueue kernels on `stream` and are re-entrant.
 *   - Return value: GG_OK (0) or a negative GG_E* code (argument errors are detected before launch;
 *     launch failures map to GG_ECUDA, cudaGetLastError text via gg_last_cuda_error()).
 *   - Boards are independent.  Rule violations cannot raise on the device, so they are reported per
 *     board in a `status` array (GG_ST_*), and a refused board's record is left unchanged - the
 *     counterpart of the reference's AssertionError (gym_go/gogame.py:59, go_env.py:54-57).
 *   - Packed record (one per board, `rec_bytes` from gg_layout, a multiple of 16):
 *       three bit-planes [black | white | invalid-for-player-to-move], each `lanes_per_board` words of
 *       `word_bits` bits; word j holds rows j*rows_per_lane ..., row r at bit (r % rows_per_lane)*(N+1),
 *       column c at +c, bit N of every row slot is a zero guard bit;
 *       then one uint32 flags word: bit0 turn (1 = white to move), bit1 previous move was a pass,
 *       bit2 game over; zero padding to the 16-byte multiple.
 *     It carries exactly the information of the reference's float64 [6,N,N] state (gogame.py:7-19).
 *   - Dense state/observation tensors are [B,6,N,N] C-contiguous with values 0/1 in `dtype`.
 */
#ifndef GYMGO_B200_H
#define GYMGO_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GG_VERSION 200 /* 0.2.0 */

#if defined(__GNUC__)
#define GG_API __attribute__((visibility("default")))
#else
#define GG_API
#endif

/* return codes */
#define GG_OK 0
#define GG_EINVAL (-1)  /* bad argument (null pointer, negative batch, unknown dtype/flag) */
#define GG_ESIZE (-2)   /* board size not supported by this build (see gg_supported) */
#define GG_EALIGN (-3)  /* a record/dense pointer is not 16-byte aligned */
#define GG_ECUDA (-4)   /* kernel launch failed; see gg_last_cuda_error() */

/* per-board status values */
#define GG_ST_OK 0
#define GG_ST_INVALID_MOVE 1 /* INVD bit set at the action (occupied, ko, suicide)  - gogame.py:59 */
#define GG_ST_OUT_OF_RANGE 2 /* action not in [0, N*N]                              - go_env.py:56-57 */
#define GG_ST_GAME_OVER 3    /* GG_STEP_REFUSE_DONE given and the board is finished - go_env.py:54 */

/* element types of dense tensors */
#define GG_U8 0
#define GG_F32 1
#define GG_F64 2 /* pack/unpack/valid_moves only (the reference's own dtype) */
#define GG_BF16 3 /* bfloat16 0/1 (0x3F80): observations, pack/unpack - feeds bf16 networks without a cast */
#define GG_F16 4  /* float16 0/1 (0x3C00): observations, pack/unpack */

/* reward_mode of gg_step / gg_rollout_step (GoEnv.reward, go_env.py:128-149) */
#define GG_REWARD_NONE 0
#define GG_REWARD_REAL 1      /* game over ? sign(black - white - komi) : 0                       (:138-139) */
#define GG_REWARD_HEURISTIC 2 /* game over ? (diff > 0 ? +N*N : -N*N) : diff = black-white-komi   (:141-147) */

/* gg_step / gg_children option bits */
#define GG_STEP_CANONICAL 1u   /* canonical=True of gogame.next_state (gogame.py:83-85, :313-321) */
#define GG_STEP_REFUSE_DONE 2u /* GoEnv.step's `assert not self.done` (go_env.py:54) */
#define GG_STEP_AUTO_RESET 4u  /* gg_step only: a board whose record is finished is first reset to gogame.init_state
                                * (GoEnv.reset, go_env.py:40-47), then plays its action - the vector-env loop in ONE launch */
#define GG_STEP_RESET_SKIPS_ACTION 8u /* with GG_STEP_AUTO_RESET: a board that was just reset does NOT play its action
                                       * this ply (status OK, empty board out) - gymnasium's next-step autoreset */

/* gg_step only: which single-ply kernel runs.  Default (neither bit): chosen from measurements like the rollout kernels -
 * thread-per-board for uint32 boards (n <= 9) in batches of >= 48 Ki when no float32 observation is written (12-16 %
 * faster there), lane-sliced otherwise.  Results are identical. */
#define GG_STEP_KERNEL_LANES 16u  /* k_step: a board spread over adjacent lanes of a warp */
#define GG_STEP_KERNEL_THREAD 32u /* k_step_tpb: one board per thread */

/* rollout kernels (gg_rollout_with); both produce identical results */
#define GG_KERNEL_AUTO (-1)    /* chosen per (n, batch) from measurements: what gg_rollout uses */
#define GG_KERNEL_LANES 0      /* k_rollout: a board spread over adjacent lanes of a warp */
#define GG_KERNEL_THREAD 1     /* k_rollout_tpb: one board per thread */

GG_API int gg_version(void);
GG_API const char *gg_last_cuda_error(void);

/* 1 if boards of side n are supported (this build: 2..19), else 0. */
GG_API int gg_supported(int n);

/* Bind the calling thread to CUDA device `ordinal` (cudaSetDevice).  The library carries its own
 * statically linked CUDA runtime; call this when the caller's current device is not 0 (the Python
 * binding does it from the tensors' device). */
GG_API int gg_set_device(int ordinal);

/* Geometry of the packed record for side n.  Any out pointer may be NULL.
 * Replaces: the state type, gogame.init_state / batch_init_state (gogame.py:22-31), govars.py:4-11. */
GG_API int gg_layout(int n, int *rec_bytes, int *lanes_per_board, int *rows_per_lane, int *word_bits);

/* dense [B,6,N,N] (dtype GG_U8/GG_F32/GG_F64) -> packed records.  turn = max of plane 2, pass = max of
 * plane 4, done = plane 5 all ones, exactly how the reference reads them (gogame.py:241-246, :200-201,
 * :208-214); stone/invalid bits are `value != 0`. */
GG_API int gg_pack(const void *dense, int dtype, int64_t batch, int n, void *rec, void *stream);

/* packed records -> dense [B,6,N,N] of dtype. */
GG_API int gg_unpack(const void *rec, int64_t batch, int n, int dtype, void *dense, void *stream);

/* Zero (= gogame.init_state) the records whose mask byte is non-zero; mask == NULL resets all.
 * Replaces: GoEnv.reset (go_env.py:40-47). */
GG_API int gg_reset(void *rec, int64_t batch, int n, const uint8_t *mask, void *stream);

/* One ply on every board: THE hot path.
 * Replaces: gogame.next_state / batch_next_states (gogame.py:34-150) with state_utils.adj_data,
 * update_pieces, compute_invalid_moves, set_turn (state_utils.py:24-83, :159-180, :214-241).
 *   rec_in, rec_out  packed records; may be the same buffer (in place)
 *   actions          int32 [B], N*N = pass
 *   status           uint8 [B] or NULL
 *   flags            GG_STEP_* bits
 *   obs_out          NULL, or dense [B,6,N,N] of obs_dtype (GG_U8, GG_F32, GG_BF16 or GG_F16) receiving the new state -
 *                    what GoEnv.step returns as the observation (go_env.py:64)
 *   done_out         NULL, or uint8 [B]: game over after this ply (gogame.game_ended, gogame.py:208-214)
 *   areas_out        NULL, or int32 [B,2] Tromp-Taylor (black, white) areas of the new state
 *                    (gogame.areas, gogame.py:275-300)
 *   reward_out       NULL, or float32 [B]: GoEnv.reward of the new state for reward_mode/komi, fused into
 *                    the kernel epilogue (areas are only flooded when the mode needs them) */
GG_API int gg_step(const void *rec_in, const int32_t *actions, void *rec_out, uint8_t *status, int64_t batch, int n,
                   uint32_t flags, void *obs_out, int obs_dtype, uint8_t *done_out, int32_t *areas_out,
                   float *reward_out, int reward_mode, float komi, void *stream);

/* Fused rollout ply: finished boards are first reset (auto-reset), then every board draws a uniformly
 * random valid action incl. pass (GoEnv.uniform_random_action, go_env.py:78-81) from Philox4x32-10
 * keyed (seed; counter = global board index board0+b, ply t) and plays it - same outputs as gg_step plus
 * the chosen actions.  Trajectories do not depend on how the batch is sharded over GPUs. */
GG_API int gg_rollout_step(void *rec, int64_t batch, int n, uint64_t seed, uint64_t board0, uint64_t t,
                           int32_t *actions_out, void *obs_out, int obs_dtype, uint8_t *done_out, int32_t *areas_out,
                           float *reward_out, int reward_mode, float komi, void *stream);

/* Device-side rollout driver: `steps` consecutive fused plies (t = t0 .. t0+steps-1, same semantics and same
 * trajectories as gg_rollout_step) executed by a PERSISTENT kernel that keeps every board in registers for
 * `plies_per_launch` plies per launch (1 = one launch per ply).  Per ply p it writes
 *   actions_log[p]  int32 [steps, B]  (may be NULL)      done_log[p]    uint8   [steps, B] (may be NULL)
 *   reward_log[p]   float32 [steps, B] (may be NULL; GoEnv.reward for reward_mode / komi)
 *   the observation of ply t into slot t % obs_ring of obs_ring_buf ([obs_ring, B, 6, N, N], may be NULL);
 * warps run ahead of each other inside a launch, so a caller that wants every observation of a launch
 * passes obs_ring >= plies_per_launch.  The packed records are read once and written once per launch. */
GG_API int gg_rollout(void *rec, int64_t batch, int n, uint64_t seed, uint64_t board0, uint64_t t0, int steps,
                      int plies_per_launch, int32_t *actions_log, void *obs_ring_buf, int obs_dtype, int obs_ring,
                      uint8_t *done_log, float *reward_log, int reward_mode, float komi, void *stream);

/* gg_rollout with an explicit kernel choice (GG_KERNEL_*) and an optional scheduling workspace.
 *   workspace    NULL, or a device buffer of >= gg_rollout_workspace_bytes(n, batch) bytes (16-byte aligned, contents
 *                irrelevant: it is cleared by a memset enqueued before every launch).  With a workspace, launches of
 *                >= 2 * block_plies plies are DYNAMICALLY SCHEDULED: the launch is cut into blocks of block_plies
 *                plies and (tile, block) work items are handed to CTAs in start order (a ticket counter), each waiting
 *                for its tile's previous block.  Boards differ persistently in cost (game phase), and a static
 *                one-CTA-per-tile launch leaves the SMs idle ~17 % of the time at the tail
 *                (profiles/r02_k_rollout_*_ncu_full.txt): +15 % (9x9) / +21 % (19x19) throughput.
 *   block_plies  plies per work item; 0 = the measured default (4).
 * Results are identical with and without a workspace.
 * gg_rollout(...) == gg_rollout_with(GG_KERNEL_AUTO, ..., NULL, 0, 0, stream). */
GG_API int gg_rollout_with(int kernel, void *rec, int64_t batch, int n, uint64_t seed, uint64_t board0, uint64_t t0,
                           int steps, int plies_per_launch, int32_t *actions_log, void *obs_ring_buf, int obs_dtype,
                           int obs_ring, uint8_t *done_log, float *reward_log, int reward_mode, float komi,
                           void *workspace, int64_t workspace_bytes, int block_plies, void *stream);
GG_API int64_t gg_rollout_workspace_bytes(int n, int64_t batch);

/* Name of the kernel GG_KERNEL_AUTO resolves to for (n, batch), and the name of kernel `kernel`. */
GG_API const char *gg_rollout_kernel(int n, int64_t batch);
GG_API const char *gg_kernel_name(int kernel);

/* Capture removal alone: on board b every group of colour 1 - player[b] that touches a point of `touch` and has no
 * empty neighbour is removed from rec[b] (in place; only that colour's plane changes).  `touch` and `killed` are
 * record arrays used as plane carriers: the BLACK plane of touch[b] holds the adjacent locations, the BLACK plane of
 * killed[b] receives the removed stones (the rest of killed[b] is zeroed).
 * Replaces: state_utils.update_pieces / batch_update_pieces (state_utils.py:159-211). */
GG_API int gg_update_pieces(void *rec, const void *touch, const int32_t *player, void *killed, int64_t batch, int n,
                            void *stream);

/* The sampler alone (no reset, no step): actions_out[b] = uniformly random valid action of board b.
 * Replaces: GoEnv.uniform_random_action / gogame.random_action (go_env.py:78-81, gogame.py:395-404). */
GG_API int gg_sample_legal(const void *rec, int64_t batch, int n, uint64_t seed, uint64_t board0, uint64_t t,
                    int32_t *actions_out, void *stream);

/* out [B, N*N+1] of dtype: 1 = valid, pass always valid.  ended_quirk != 0 reproduces the single-board
 * gogame.valid_moves (all ones once the game has ended, gogame.py:153-161); 0 gives batch_valid_moves
 * (gogame.py:164-172).  invalid_moves = 1 - this. */
GG_API int gg_valid_moves(const void *rec, int64_t batch, int n, int ended_quirk, int dtype, void *out, void *stream);

/* Full legal-move expansion: slot (b, a) = next_state(parent b, action a) for every valid a, zeros
 * otherwise (padded=True of gogame.children, gogame.py:175-186).
 *   child_rec   NULL or packed [B, A] records          valid   NULL or uint8 [B, A]
 *   child_obs   NULL or dense [B, A, 6, N, N]           status  NULL or uint8 [B]: 1 if a "valid" action was
 *               of obs_dtype (U8/F32/BF16/F16)                  refused (finished parent with stones: the
 *                                                               reference asserts there, gogame.py:117)
 *   flags       GG_STEP_CANONICAL or 0 */
GG_API int gg_children(const void *rec, int64_t batch, int n, uint32_t flags, void *child_rec, void *child_obs,
                int obs_dtype, uint8_t *valid, uint8_t *status, void *stream);

/* int32 [B,2] (black area, white area).  Replaces: gogame.areas / batch_areas (gogame.py:275-310);
 * winning = sign(black - white - komi) is left to the caller (gogame.py:225-238). */
GG_API int gg_areas(const void *rec, int64_t batch, int n, int32_t *out, void *stream);

/* Canonical form: boards with white to move get their stone planes swapped and turn cleared.
 * Replaces: gogame.canonical_form / batch_canonical_form (gogame.py:313-337). */
GG_API int gg_canonical(const void *rec_in, void *rec_out, int64_t batch, int n, void *stream);

/* One of the 8 dihedral transforms applied to packed records (stone and invalid planes; flags copied), out of
 * place.  `sym` (0..7) is the index into gogame.all_symmetries (gogame.py:358-382): bit 0 mirrors the columns
 * (np.flip axis 2), bit 1 mirrors the rows (np.flip axis 1), bit 2 turns the result a quarter counter-clockwise
 * (np.rot90 over axes 1,2), applied in that order.  Pinned by tests/golden/misc.npz (outputs of the reference).
 * Replaces: random_symmetry / all_symmetries (gogame.py:340-382). */
GG_API int gg_symmetry(const void *rec_in, void *rec_out, int64_t batch, int n, int sym, void *stream);

/* Measurement aid: fill `bytes` (a multiple of 16) of device memory with 16-byte streaming stores, every warp
 * streaming contiguous runs of `run_bytes` (a multiple of 512) - the store instruction and pattern of the observation
 * emission and nothing else.  bench.py times it to report the pure-write ceiling of the GPU next to the copy-based
 * HBM peak of MEASURED_PEAKS.json. */
GG_API int gg_probe_write(void *buf, int64_t bytes, int64_t run_bytes, void *stream);

/* HOST codec (the only entry point that takes host pointers and runs on the CPU): packed records in host memory ->
 * dense [B,6,N,N] of dtype in host memory, `threads` worker threads (<= 0: all hardware threads, capped at 64) taken
 * from a persistent pool.  For consumers that fetch the packed records over PCIe instead of the 40x larger dense
 * observation (BatchedGoEnv.host_stepper(transport="packed")): with AVX-512 (mask moves + non-temporal stores, chosen
 * at run time) the expansion runs at host-memory write speed, which beats the PCIe transfer of the dense tensor.
 * A 64-byte aligned `dense_host` gets the streaming-store path.  Calls are serialised inside the library.
 * No Go rules run here - it is gg_unpack's layout, nothing else. */
GG_API int gg_host_unpack(const void *rec_host, int64_t batch, int n, int dtype, void *dense_host, int threads);
/* which implementation gg_host_unpack runs on this CPU ("avx512 ..." or "scalar") */
GG_API const char *gg_host_unpack_path(void);

#ifdef __cplusplus
}
#endif
#endif /* GYMGO_B200_H */
