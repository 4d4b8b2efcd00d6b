// gg_api.cu - the C ABI of include/gymgo_b200.h: argument checks + dispatch on the board size.
// No allocation, no synchronisation, no global state besides the last-error string.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/gymgo_b200.h"
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

// Which persistent rollout kernel serves (n, batch).  Measured (profiles/r01_variant_threshold.json, f32
// observations): both kernels sit within a few percent of the pure-write floor; thread-per-board is ahead by up
// to 5 % on small boards around 64 Ki boards (the lane-sliced kernel cannot fully overlap rules and stores there),
// behind below 32 Ki (too few warps) and level or slightly behind at 128 Ki (both write-bound).
// GG_ROLLOUT_VARIANT=0/1/2 overrides the choice for A/B measurements.
static int rollout_variant(const SizeVTable* v, int64_t batch) {
    const char* forced = getenv("GG_ROLLOUT_VARIANT");
    if (forced) return atoi(forced);
    const bool tpb_ok = v->wordbits == 32 && v->lpb <= 3;
    return (tpb_ok && batch >= 49152 && batch < 98304) ? 1 : 0;
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
    if (batch < 0 || (flags & ~(GG_STEP_CANONICAL | GG_STEP_REFUSE_DONE))) return GG_EINVAL;
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

GG_API int gg_rollout(void* rec, int64_t batch, int n, uint64_t seed, uint64_t board0, uint64_t t0, int steps,
               int plies_per_launch, int32_t* actions_log, void* obs_ring_buf, int obs_dtype, int obs_ring,
               uint8_t* done_log, float* reward_log, int reward_mode, float komi, void* stream) {
    const SizeVTable* v = lookup(n);
    if (!v) return GG_ESIZE;
    if (batch < 0 || steps < 0 || plies_per_launch < 1 || (batch > 0 && !rec)) return GG_EINVAL;
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
    // developer switch (A/B measurements only): GG_ROLLOUT_VARIANT=1 selects the thread-per-board kernel on small boards
    a.variant = rollout_variant(v, batch);
    const char* slice_k = getenv("GG_ROLLOUT_K");
    a.slice_k = slice_k ? atoi(slice_k) : 2;
    for (int p = 0; p < steps; p += plies_per_launch) {
        a.t0 = t0 + uint64_t(p);
        a.plies = steps - p < plies_per_launch ? steps - p : plies_per_launch;
        a.actions_log = actions_log ? actions_log + size_t(p) * size_t(batch) : nullptr;
        a.done_log = done_log ? done_log + size_t(p) * size_t(batch) : nullptr;
        a.reward_log = (reward_log && reward_mode != GG_REWARD_NONE) ? reward_log + size_t(p) * size_t(batch) : nullptr;
        cudaError_t e = v->rollout(a, static_cast<cudaStream_t>(stream));
        if (e != cudaSuccess) return finish(e);
    }
    return GG_OK;
}

GG_API const char* gg_rollout_kernel(int n, int64_t batch) {
    const SizeVTable* v = lookup(n);
    if (!v) return "";
    const int variant = rollout_variant(v, batch);
    return variant == 1 ? "k_rollout_tpb (thread per board)"
                        : (variant == 2 ? "k_rollout_sliced (several words per lane)" : "k_rollout (lane-sliced boards)");
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
