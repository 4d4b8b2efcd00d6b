// gg_api.cu - the C ABI of include/gymgo_b200.h: argument checks + dispatch on the board size.
// No allocation, no synchronisation, no global state besides the last-error string.
#include <stdio.h>
#include <string.h>

#include <thread>

#include "../../include/gymgo_b200.h"
#include "gg_host.h"
#include "gg_kernels.cuh"

namespace gg {
#define GG_SIZES(X) X(2) X(3) X(4) X(5) X(6) X(7) X(8) X(9) X(10) X(11) X(12) X(13) X(14) X(15) X(16) X(17) X(18) X(19)
#define X(N) extern const SizeVTable vtable_n##N;
GG_SIZES(X)
#undef X

static const SizeVTable* lookup(int n) {
    switch (n) {
#define X(N) \
    case N:  \
        return &vtable_n##N;
        GG_SIZES(X)
#undef X
    }
    return nullptr;
}

static thread_local char g_err[256] = "";

// Which persistent rollout kernel serves (n, batch) when the caller asks for GG_KERNEL_AUTO.  Measured on the B200
// with f32 observations (profiles/r01_variant_threshold.json, profiles/r02_kernel_ab.json): both kernels sit within a
// few percent of the pure-write floor on small boards; thread-per-board is ahead by up to 5 % on small boards around
// 64 Ki boards, behind below 32 Ki (too few warps) and level at 128 Ki (write-bound either way).
static int auto_kernel(const SizeVTable* v, int64_t batch) {
    const bool tpb_ok = v->wordbits == 32 && v->lpb <= 3;
    if (tpb_ok && batch >= 49152 && batch < 98304) return GG_KERNEL_THREAD;
    return GG_KERNEL_LANES;
}
static const char* kernel_name(int k) {
    return k == GG_KERNEL_THREAD ? "k_rollout_tpb (thread per board)" : "k_rollout (lane-sliced boards)";
}

static int finish(cudaError_t e) {
    if (e == cudaSuccess) return GG_OK;
    snprintf(g_err, sizeof g_err, "%s: %s", cudaGetErrorName(e), cudaGetErrorString(e));
    return GG_ECUDA;
}
static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
static bool dense_dtype_ok(int dt, bool allow_f64) {
    return dt == GG_U8 || dt == GG_F32 || dt == GG_BF16 || dt == GG_F16 || (allow_f64 && dt == GG_F64);
}
}  // namespace gg

namespace gg {
// Measurement aid (gg_probe_write): nothing but 16-byte streaming stores, the same instruction (st.global.cs.v4) and
// access pattern the observation emission uses - the pure-write ceiling of the GPU, which bench.py reports next to the
// copy-based HBM peak (a copy alternates reads and writes on the DRAM bus and tops out lower than a write stream).
static __global__ void __launch_bounds__(256) k_probe_write(float4* buf, long long vectors, long long chunk) {
    // every warp streams contiguous runs of `chunk` vectors (like a rollout warp streams its boards' observations)
    const float4 v = make_float4(1.f, 0.f, 1.f, 0.f);
    const long long warps = (long long)gridDim.x * (blockDim.x >> 5);
    const long long warp = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    for (long long base = warp * chunk; base < vectors; base += warps * chunk) {
        const long long end = base + chunk < vectors ? base + chunk : vectors;
#pragma unroll 4
        for (long long i = base + lane; i < end; i += 32) __stcs(buf + i, v);
    }
}
}  // namespace gg

using namespace gg;

extern "C" {

GG_API int gg_version(void) { return GG_VERSION; }
GG_API const char* gg_last_cuda_error(void) { return g_err; }
GG_API int gg_supported(int n) { return lookup(n) != nullptr; }
GG_API int gg_set_device(int ordinal) { return finish(cudaSetDevice(ordinal)); }

GG_API int gg_layout(int n, int* rec_bytes, int* lanes_per_board, int* rows_per_lane, int* word_bits) {
    const SizeVTable* v = lookup(n);
    if (!v) return GG_ESIZE;
    if (rec_bytes) *rec_bytes = v->rec_bytes;
    if (lanes_per_board) *lanes_per_board = v->lpb;
    if (rows_per_lane) *rows_per_lane = v->rpl;
    if (word_bits) *word_bits = v->wordbits;
    return GG_OK;
}

GG_API int gg_pack(const void* dense, int dtype, int64_t batch, int n, void* rec, void* stream) {
    const SizeVTable* v = lookup(n);
    if (!v) return GG_ESIZE;
    if (batch < 0 || !dense_dtype_ok(dtype, true) || (batch > 0 && (!dense || !rec))) return GG_EINVAL;
    if (!aligned16(rec)) return GG_EALIGN;
    return finish(v->pack(dense, dtype, batch, static_cast<uint32_t*>(rec), static_cast<cudaStream_t>(stream)));
}

GG_API int gg_unpack(const void* rec, int64_t batch, int n, int dtype, void* dense, void* stream) {
    const SizeVTable* v = lookup(n);
    if (!v) return GG_ESIZE;
    if (batch < 0 || !dense_dtype_ok(dtype, true) || (batch > 0 && (!dense || !rec))) return GG_EINVAL;
    if (!aligned16(rec)) return GG_EALIGN;
    return finish(v->unpack(static_cast<const uint32_t*>(rec), batch, dtype, dense, static_cast<cudaStream_t>(stream)));
}

GG_API int gg_reset(void* rec, int64_t batch, int n, const uint8_t* mask, void* stream) {
    const SizeVTable* v = lookup(n);
    if (!v) return GG_ESIZE;
    if (batch < 0 || (batch > 0 && !rec)) return GG_EINVAL;
    if (!aligned16(rec)) return GG_EALIGN;
    return finish(v->reset(static_cast<uint32_t*>(rec), batch, mask, static_cast<cudaStream_t>(stream)));
}

GG_API int gg_step(const void* rec_in, const int32_t* actions, void* rec_out, uint8_t* status, int64_t batch, int n,
            uint32_t flags, void* obs_out, int obs_dtype, uint8_t* done_out, int32_t* areas_out, float* reward_out,
            int reward_mode, float komi, void* stream) {
    const SizeVTable* v = lookup(n);
    if (!v) return GG_ESIZE;
    if (batch < 0 || (flags & ~(GG_STEP_CANONICAL | GG_STEP_REFUSE_DONE | GG_STEP_AUTO_RESET | GG_STEP_RESET_SKIPS_ACTION |
                                GG_STEP_KERNEL_LANES | GG_STEP_KERNEL_THREAD))) return GG_EINVAL;
    if ((flags & GG_STEP_KERNEL_LANES) && (flags & GG_STEP_KERNEL_THREAD)) return GG_EINVAL;
    if (batch > 0 && (!rec_in || !rec_out || !actions)) return GG_EINVAL;
    if (obs_out && !dense_dtype_ok(obs_dtype, false)) return GG_EINVAL;
    if (reward_mode < GG_REWARD_NONE || reward_mode > GG_REWARD_HEURISTIC) return GG_EINVAL;
    if (!aligned16(rec_in) || !aligned16(rec_out) || !aligned16(obs_out)) return GG_EALIGN;
    StepArgs a;
    memset(&a, 0, sizeof a);
    a.rec_in = static_cast<const uint32_t*>(rec_in);
    a.rec_out = static_cast<uint32_t*>(rec_out);
    a.actions_in = actions;
    a.status = status;
    a.obs = obs_out;
    a.obs_dtype = obs_dtype;
    a.done_out = done_out;
    a.areas_out = areas_out;
    a.reward_out = reward_mode == GG_REWARD_NONE ? nullptr : reward_out;
    a.reward_mode = reward_mode;
    a.komi = komi;
    a.slots = batch;
    a.opts = flags;
    return finish(v->step(a, MODE_STEP, static_cast<cudaStream_t>(stream)));
}

GG_API int gg_rollout_step(void* rec, int64_t batch, int n, uint64_t seed, uint64_t board0, uint64_t t, int32_t* actions_out,
                    void* obs_out, int obs_dtype, uint8_t* done_out, int32_t* areas_out, float* reward_out,
                    int reward_mode, float komi, void* stream) {
    const SizeVTable* v = lookup(n);
    if (!v) return GG_ESIZE;
    if (batch < 0 || (batch > 0 && !rec)) return GG_EINVAL;
    if (obs_out && !dense_dtype_ok(obs_dtype, false)) return GG_EINVAL;
    if (reward_mode < GG_REWARD_NONE || reward_mode > GG_REWARD_HEURISTIC) return GG_EINVAL;
    if (!aligned16(rec) || !aligned16(obs_out)) return GG_EALIGN;
    StepArgs a;
    memset(&a, 0, sizeof a);
    a.rec_in = static_cast<const uint32_t*>(rec);
    a.rec_out = static_cast<uint32_t*>(rec);
    a.actions_out = actions_out;
    a.obs = obs_out;
    a.obs_dtype = obs_dtype;
    a.done_out = done_out;
    a.areas_out = areas_out;
    a.reward_out = reward_mode == GG_REWARD_NONE ? nullptr : reward_out;
    a.reward_mode = reward_mode;
    a.komi = komi;
    a.slots = batch;
    a.seed = seed;
    a.board0 = board0;
    a.t = t;
    return finish(v->step(a, MODE_ROLLOUT, static_cast<cudaStream_t>(stream)));
}

// workspace of the dynamically scheduled rollout: one ticket counter + one progress word per tile (tiles hold >= 8 boards)
static int64_t rollout_workspace_bytes(int64_t batch) { return ((batch / 8 + 2) * 4 + 15) / 16 * 16; }
// plies per scheduling block when the caller passes 0: measured best 3..5 on both headline configs
// (tools/sweep_block_plies.py, profiles/r02_block_plies_sweep.json)
static const int kDefaultBlockPlies = 4;

GG_API int64_t gg_rollout_workspace_bytes(int n, int64_t batch) {
    return (lookup(n) && batch >= 0) ? rollout_workspace_bytes(batch) : GG_ESIZE;
}
GG_API int gg_rollout_with(int kernel, void* rec, int64_t batch, int n, uint64_t seed, uint64_t board0, uint64_t t0, int steps,
                    int plies_per_launch, int32_t* actions_log, void* obs_ring_buf, int obs_dtype, int obs_ring,
                    uint8_t* done_log, float* reward_log, int reward_mode, float komi, void* workspace,
                    int64_t workspace_bytes, int block_plies, void* stream) {
    const SizeVTable* v = lookup(n);
    if (!v) return GG_ESIZE;
    if (batch < 0 || steps < 0 || plies_per_launch < 1 || (batch > 0 && !rec)) return GG_EINVAL;
    if (kernel < GG_KERNEL_AUTO || kernel > GG_KERNEL_THREAD) return GG_EINVAL;
    if (obs_ring_buf && (!dense_dtype_ok(obs_dtype, false) || obs_ring < 1)) return GG_EINVAL;
    if (reward_mode < GG_REWARD_NONE || reward_mode > GG_REWARD_HEURISTIC) return GG_EINVAL;
    if (!aligned16(rec) || !aligned16(obs_ring_buf)) return GG_EALIGN;
    RolloutArgs a;
    memset(&a, 0, sizeof a);
    a.rec = static_cast<uint32_t*>(rec);
    a.boards = batch;
    a.seed = seed;
    a.board0 = board0;
    a.reward_mode = reward_mode;
    a.komi = komi;
    a.obs_ring = obs_ring_buf;
    a.obs_dtype = obs_dtype;
    a.ring = obs_ring_buf ? obs_ring : 1;
    a.variant = kernel == GG_KERNEL_AUTO ? auto_kernel(v, batch) : kernel;
    if (workspace && (workspace_bytes < rollout_workspace_bytes(batch) || !aligned16(workspace))) return GG_EINVAL;
    if (block_plies < 0) return GG_EINVAL;
    a.block_plies = block_plies > 0 ? block_plies : kDefaultBlockPlies;
    for (int p = 0; p < steps; p += plies_per_launch) {
        a.t0 = t0 + uint64_t(p);
        a.plies = steps - p < plies_per_launch ? steps - p : plies_per_launch;
        a.ws = nullptr;
        if (workspace && a.plies >= 2 * a.block_plies) {           // long launches are scheduled dynamically
            cudaError_t e = cudaMemsetAsync(workspace, 0, size_t(rollout_workspace_bytes(batch)), static_cast<cudaStream_t>(stream));
            if (e != cudaSuccess) return finish(e);
            a.ws = static_cast<int*>(workspace);
        }
        a.actions_log = actions_log ? actions_log + size_t(p) * size_t(batch) : nullptr;
        a.done_log = done_log ? done_log + size_t(p) * size_t(batch) : nullptr;
        a.reward_log = (reward_log && reward_mode != GG_REWARD_NONE) ? reward_log + size_t(p) * size_t(batch) : nullptr;
        cudaError_t e = v->rollout(a, static_cast<cudaStream_t>(stream));
        if (e != cudaSuccess) return finish(e);
    }
    return GG_OK;
}

GG_API int gg_rollout(void* rec, int64_t batch, int n, uint64_t seed, uint64_t board0, uint64_t t0, int steps,
               int plies_per_launch, int32_t* actions_log, void* obs_ring_buf, int obs_dtype, int obs_ring,
               uint8_t* done_log, float* reward_log, int reward_mode, float komi, void* stream) {
    return gg_rollout_with(GG_KERNEL_AUTO, rec, batch, n, seed, board0, t0, steps, plies_per_launch, actions_log, obs_ring_buf,
                           obs_dtype, obs_ring, done_log, reward_log, reward_mode, komi, nullptr, 0, 0, stream);
}

GG_API const char* gg_rollout_kernel(int n, int64_t batch) {
    const SizeVTable* v = lookup(n);
    return v ? kernel_name(auto_kernel(v, batch)) : "";
}

GG_API const char* gg_kernel_name(int kernel) {
    return (kernel >= GG_KERNEL_LANES && kernel <= GG_KERNEL_THREAD) ? kernel_name(kernel) : "";
}

GG_API int gg_update_pieces(void* rec, const void* touch, const int32_t* player, void* killed, int64_t batch, int n, void* stream) {
    const SizeVTable* v = lookup(n);
    if (!v) return GG_ESIZE;
    if (batch < 0 || (batch > 0 && (!rec || !touch || !player || !killed))) return GG_EINVAL;
    if (rec == killed || touch == killed) return GG_EINVAL;
    if (!aligned16(rec) || !aligned16(touch) || !aligned16(killed)) return GG_EALIGN;
    return finish(v->capture(static_cast<uint32_t*>(rec), static_cast<const uint32_t*>(touch), player,
                             static_cast<uint32_t*>(killed), batch, static_cast<cudaStream_t>(stream)));
}

GG_API int gg_host_unpack(const void* rec_host, int64_t batch, int n, int dtype, void* dense_host, int threads) {
    const SizeVTable* v = lookup(n);
    if (!v) return GG_ESIZE;
    if (batch < 0 || !dense_dtype_ok(dtype, true) || (batch > 0 && (!rec_host || !dense_host))) return GG_EINVAL;
    if (threads <= 0) threads = int(std::thread::hardware_concurrency());
    if (threads > 64) threads = 64;
    host_unpack(static_cast<const uint8_t*>(rec_host), batch, n, v->lpb, v->rpl, v->wordbits, v->rec_bytes, dtype, dense_host,
                threads < 1 ? 1 : threads);
    return GG_OK;
}
GG_API const char* gg_host_unpack_path(void) { return host_unpack_path(); }

GG_API int gg_probe_write(void* buf, int64_t bytes, int64_t run_bytes, void* stream) {
    if (bytes < 0 || (bytes > 0 && !buf) || (bytes & 15) || run_bytes < 512 || (run_bytes & 511)) return GG_EINVAL;
    if (!aligned16(buf)) return GG_EALIGN;
    if (bytes == 0) return GG_OK;
    k_probe_write<<<148 * 8, 256, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<float4*>(buf), bytes / 16, run_bytes / 16);
    return finish(cudaGetLastError());
}

GG_API int gg_sample_legal(const void* rec, int64_t batch, int n, uint64_t seed, uint64_t board0, uint64_t t,
                    int32_t* actions_out, void* stream) {
    const SizeVTable* v = lookup(n);
    if (!v) return GG_ESIZE;
    if (batch < 0 || (batch > 0 && (!rec || !actions_out))) return GG_EINVAL;
    if (!aligned16(rec)) return GG_EALIGN;
    return finish(v->sample(static_cast<const uint32_t*>(rec), batch, seed, board0, t, actions_out,
                            static_cast<cudaStream_t>(stream)));
}

GG_API int gg_valid_moves(const void* rec, int64_t batch, int n, int ended_quirk, int dtype, void* out, void* stream) {
    const SizeVTable* v = lookup(n);
    if (!v) return GG_ESIZE;
    if (batch < 0 || !(dtype == GG_U8 || dtype == GG_F32 || dtype == GG_F64) || (batch > 0 && (!rec || !out))) return GG_EINVAL;
    if (!aligned16(rec)) return GG_EALIGN;
    return finish(v->valid(static_cast<const uint32_t*>(rec), batch, ended_quirk, dtype, out, static_cast<cudaStream_t>(stream)));
}

GG_API int gg_children(const void* rec, int64_t batch, int n, uint32_t flags, void* child_rec, void* child_obs, int obs_dtype,
                uint8_t* valid, uint8_t* status, void* stream) {
    const SizeVTable* v = lookup(n);
    if (!v) return GG_ESIZE;
    if (batch < 0 || (flags & ~GG_STEP_CANONICAL) || (batch > 0 && !rec)) return GG_EINVAL;
    if (child_obs && !dense_dtype_ok(obs_dtype, false)) return GG_EINVAL;
    if (!aligned16(rec) || !aligned16(child_rec) || !aligned16(child_obs)) return GG_EALIGN;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (status && batch > 0) {
        cudaError_t e = cudaMemsetAsync(status, 0, size_t(batch), s);
        if (e != cudaSuccess) return finish(e);
    }
    StepArgs a;
    memset(&a, 0, sizeof a);
    a.rec_in = static_cast<const uint32_t*>(rec);
    a.rec_out = static_cast<uint32_t*>(child_rec);
    a.status = status;
    a.valid_out = valid;
    a.obs = child_obs;
    a.obs_dtype = obs_dtype;
    a.slots = batch * (int64_t(n) * n + 1);
    a.opts = flags;
    return finish(v->step(a, MODE_CHILDREN, s));
}

GG_API int gg_areas(const void* rec, int64_t batch, int n, int32_t* out, void* stream) {
    const SizeVTable* v = lookup(n);
    if (!v) return GG_ESIZE;
    if (batch < 0 || (batch > 0 && (!rec || !out))) return GG_EINVAL;
    if (!aligned16(rec)) return GG_EALIGN;
    return finish(v->areas(static_cast<const uint32_t*>(rec), batch, out, static_cast<cudaStream_t>(stream)));
}

GG_API int gg_canonical(const void* rec_in, void* rec_out, int64_t batch, int n, void* stream) {
    const SizeVTable* v = lookup(n);
    if (!v) return GG_ESIZE;
    if (batch < 0 || (batch > 0 && (!rec_in || !rec_out))) return GG_EINVAL;
    if (!aligned16(rec_in) || !aligned16(rec_out)) return GG_EALIGN;
    return finish(v->canonical(static_cast<const uint32_t*>(rec_in), static_cast<uint32_t*>(rec_out), batch,
                               static_cast<cudaStream_t>(stream)));
}

GG_API int gg_symmetry(const void* rec_in, void* rec_out, int64_t batch, int n, int sym, void* stream) {
    const SizeVTable* v = lookup(n);
    if (!v) return GG_ESIZE;
    if (batch < 0 || sym < 0 || sym > 7 || (batch > 0 && (!rec_in || !rec_out || rec_in == rec_out))) return GG_EINVAL;
    if (!aligned16(rec_in) || !aligned16(rec_out)) return GG_EALIGN;
    return finish(v->symmetry(static_cast<const uint32_t*>(rec_in), static_cast<uint32_t*>(rec_out), batch, sym,
                              static_cast<cudaStream_t>(stream)));
}

}  // extern "C"
