// gg_kernels.cuh - sm_100a kernels of the batched Go engine (templates over the board geometry).
//
// Work decomposition: a board occupies LPB adjacent lanes of a warp (9x9: 3 lanes x
// 3 rows in a uint32, 10 boards per warp; 19x19: 7 lanes x 3 rows in a uint64, 4 boards per warp), a CTA
// owns a contiguous tile of BT boards:
//   1. the tile's packed records are staged into shared memory with ONE bulk async copy (TMA,
//      cp.async.bulk + mbarrier; SASS: UBLKCP) - coalesced by construction;
//   2. every lane keeps its slice of the three bit-planes in registers and runs gg::Algo (floods are
//      carry chains + neighbour shuffles, votes are ballots) - no shared-memory traffic in the rules;
//   3. new records go back to shared memory and leave with one bulk async store;
//   4. the dense [6,N,N] observation (what GoEnv.step returns) is produced from a bit stream in shared memory
//      (per tile in k_step, per warp in the rollout kernels; assembled with atomicOr by the lane-sliced
//      kernels, in registers + funnel shifts + one shuffle by the thread-per-board kernel): each 16-byte store
//      expands one nibble (f32), one byte (bf16/f16) or 16 bits (u8) through a small table - fully coalesced
//      128-bit streaming stores, which is the HBM traffic that dominates the step.
// The persistent rollout kernels (k_rollout, k_rollout_tpb) keep the boards in registers over the plies of a work
// item and are dynamically scheduled: (tile, 4-ply block) tickets, see RolloutArgs.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "gg_algo.cuh"
#include "gg_array_ops.cuh"

namespace gg {

enum { MODE_STEP = 0, MODE_ROLLOUT = 1, MODE_CHILDREN = 2 };
enum { DT_U8 = 0, DT_F32 = 1, DT_F64 = 2, DT_BF16 = 3, DT_F16 = 4 };

struct StepArgs {
    const uint32_t* rec_in;     // STEP/ROLLOUT: [B] records; CHILDREN: [B] parent records
    uint32_t* rec_out;          // STEP/ROLLOUT: [B]; CHILDREN: [B*A] child records or NULL
    const int32_t* actions_in;  // STEP
    int32_t* actions_out;       // ROLLOUT
    uint8_t* status;            // STEP: [B]; CHILDREN: [B] parent status
    uint8_t* valid_out;         // CHILDREN: [B*A] or NULL
    void* obs;                  // dense output or NULL
    int obs_dtype;              // DT_U8 / DT_F32
    uint8_t* done_out;          // [slots] or NULL
    int32_t* areas_out;         // [slots][2] or NULL
    float* reward_out;          // [slots] or NULL (GoEnv.reward epilogue)
    int reward_mode;            // 1 real, 2 heuristic
    float komi;
    long long slots;            // boards (STEP/ROLLOUT) or B*A (CHILDREN)
    uint32_t opts;
    unsigned long long seed, board0, t;
};

// ------------------------------------------------------------------------------------ PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "GG_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra GG_DONE;\n"
        "bra GG_WAIT;\n"
        "GG_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// gpu-scope acquire / release on a global flag, and the generic <-> async proxy fence that orders them against TMA
// traffic to global memory (records written by one CTA's bulk store are read by another CTA's bulk load)
__device__ __forceinline__ int ld_acquire_gpu(const int* p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_gpu(int* p, int v) {
    asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
// TMA bulk copy global -> shared, completion counted in bytes on the mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void bulk_g2s(void* sdst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(sdst)),
                 "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// TMA bulk copy shared -> global
__device__ __forceinline__ void bulk_s2g(void* gdst, const void* ssrc, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(ssrc)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit_wait_all() {
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

// ------------------------------------------------------------------- device Ops policy for gg::Algo
template <class G>
struct DevOps {
    typedef typename G::W W;
    typedef W P;
    static constexpr unsigned FULL = 0xffffffffu;
    static constexpr unsigned GM = G::LPB >= 32 ? 0xffffffffu : ((1u << G::LPB) - 1u);
    int j;            // my lane inside the board's lane group
    unsigned gshift;  // first warp lane of my board's group
    W fullmask;       // real points of my slice (0 for ghost lanes / boards past the end)

    __device__ __forceinline__ void init(int lane, bool real) {
        const int slot = lane / G::LPB;
        const bool ghost = slot >= G::BPW;
        gshift = ghost ? unsigned(G::BPW * G::LPB) : unsigned(slot * G::LPB);
        j = lane - int(gshift);
        fullmask = (!ghost && real) ? G::rows_mask(G::rows_in_lane(j)) : W(0);
    }
    __device__ __forceinline__ P zero() const { return W(0); }
    __device__ __forceinline__ P full() const { return fullmask; }
    __device__ __forceinline__ P andnot(P a, P b) const { return a & ~b; }
    __device__ __forceinline__ P east(P x) const { return x << 1; }
    __device__ __forceinline__ P west(P x) const { return x >> 1; }
    __device__ __forceinline__ P south(P x) const {   // result[r] = x[r-1]
        W r = 0;
        if (G::RPL > 1) r = x << (G::S % G::WB);
        if (G::LPB > 1) {
            const W prev = __shfl_up_sync(FULL, x, 1);
            if (j > 0) r |= prev >> ((G::RPL - 1) * G::S);
        }
        return r;
    }
    __device__ __forceinline__ P north(P x) const {   // result[r] = x[r+1]
        W r = 0;
        if (G::RPL > 1) r = x >> (G::S % G::WB);
        if (G::LPB > 1) {
            const W next = __shfl_down_sync(FULL, x, 1);
            if (j < G::LPB - 1) r |= next << ((G::RPL - 1) * G::S);   // (its rows 1.. land above the RPL row slots)
        }
        return r;
    }
    __device__ __forceinline__ P rev(P x) const { return w_rev(x); }
    __device__ __forceinline__ P hfill(P s, P m, P mrev) const { return w_hfill(s, m, mrev); }
    __device__ __forceinline__ bool any(P x) const { return __any_sync(FULL, x != 0); }
    __device__ __forceinline__ unsigned group_ballot(bool pred) const {
        return (__ballot_sync(FULL, pred) >> gshift) & GM;
    }
    __device__ __forceinline__ bool any_board(P x) const {
        if (G::LPB == 1) return x != 0;
        return group_ballot(x != 0) != 0;
    }
    __device__ __forceinline__ int count2(P x) const {
        const int c = w_popc(x);
        if (G::LPB == 1) return c > 2 ? 2 : c;
        const unsigned b1 = group_ballot(c >= 1), b2 = group_ballot(c >= 2);
        return (b2 || __popc(b1) >= 2) ? 2 : (b1 ? 1 : 0);
    }
    __device__ __forceinline__ int popc(P x) const {
        const int c = w_popc(x);
        if (G::LPB == 1) return c;
        int total = 0;
#pragma unroll
        for (int l = 0; l < G::LPB; ++l) total += __shfl_sync(FULL, c, int(gshift) + l);
        return total;
    }
    __device__ __forceinline__ P lowest(P x) const {
        const W low = x & (~x + 1);
        if (G::LPB == 1) return low;
        const unsigned nz = group_ballot(x != 0);
        return (j == __ffs(int(nz)) - 1) ? low : W(0);
    }
    __device__ __forceinline__ P single(int pt) const {
        const int r = pt / G::N, c = pt - r * G::N;
        const int lj = r / G::RPL;
        return lj == j ? W(1) << ((r - lj * G::RPL) * G::S + c) : W(0);
    }
    __device__ __forceinline__ P pick(bool c, P a, P b) const { return c ? a : b; }
    __device__ __forceinline__ int kth_point(P x, int k) const {
        const int c = w_popc(x);
        int before = 0;
        if (G::LPB > 1) {
#pragma unroll
            for (int l = 0; l < G::LPB; ++l) {
                const int cl = __shfl_sync(FULL, c, int(gshift) + l);
                if (l < j) before += cl;
            }
        }
        const int kl = k - before;
        const bool mine = kl >= 0 && kl < c;
        int pt = -1;
        if (mine) {
            const int bit = w_select(x, kl);
            const int row = bit / G::S;
            pt = (j * G::RPL + row) * G::N + (bit - row * G::S);
        }
        if (G::LPB > 1) {
            const unsigned owner = group_ballot(mine);
            pt = __shfl_sync(FULL, pt, int(gshift) + (owner ? __ffs(int(owner)) - 1 : 0));
        }
        return pt;
    }
};

// -------------------------------------------------------------------------- record word access
template <class G>
__device__ __forceinline__ typename G::W rec_word(const uint32_t* rec, int plane, int j) {
    const uint32_t* p = rec + (plane * G::LPB + j) * G::WW;
    if (G::WW == 1) return typename G::W(p[0]);
    return typename G::W(*reinterpret_cast<const uint64_t*>(p));
}
template <class G>
__device__ __forceinline__ void rec_word_store(uint32_t* rec, int plane, int j, typename G::W v) {
    uint32_t* p = rec + (plane * G::LPB + j) * G::WW;
    if (G::WW == 1) p[0] = uint32_t(v);
    else *reinterpret_cast<uint64_t*>(p) = uint64_t(v);
}

// CTA shape: BT (boards per tile) must be a multiple of 8 so that every tile of the dense output starts
// 16-byte aligned for both f32 and u8 (6*N*N is even).
template <class G>
struct Tile {
    static constexpr int WPC = (G::BPW % 2 == 0) ? 4 : 8;
    static constexpr int THREADS = WPC * 32;
    static constexpr int BT = WPC * G::BPW;
    static constexpr int DENSE = 6 * G::NP;                      // elements per board
    static constexpr int STREAM_W32 = (BT * DENSE + 31) / 32 + 2;
    // persistent rollout kernel: keep every CTA of the headline batches resident (9x9: 11.1 CTAs per SM)
    static constexpr int ROLLOUT_MIN_BLOCKS = G::WB == 32 ? 12 : 8;
    static_assert(BT % 8 == 0, "tile must keep the dense output 16-byte aligned");
};

// --------------------------------------------------------------------- dense-observation bit stream
__device__ __forceinline__ void stream_put32(uint32_t* s, int off, uint32_t v) {
    if (v == 0) return;
    const int w = off >> 5, sh = off & 31;
    atomicOr(&s[w], v << sh);
    if (sh != 0 && (v >> (32 - sh)) != 0) atomicOr(&s[w + 1], v >> (32 - sh));
}
__device__ __forceinline__ void stream_put(uint32_t* s, int off, uint32_t v) { stream_put32(s, off, v); }
__device__ __forceinline__ void stream_put(uint32_t* s, int off, uint64_t v) {
    stream_put32(s, off, uint32_t(v));
    stream_put32(s, off + 32, uint32_t(v >> 32));
}
// drop the guard bits of my slice: RPL rows of N bits, row-major
template <class G>
__device__ __forceinline__ typename G::W compact_rows(typename G::W w) {
    typename G::W out = 0;
#pragma unroll
    for (int i = 0; i < G::RPL; ++i) out |= ((w >> (i * G::S)) & G::row_bits()) << (i * G::N);
    return out;
}
template <class G>
__device__ __forceinline__ void stream_put_board(uint32_t* s_bits, int board_bit0, int j, typename G::W black,
                                                 typename G::W white, typename G::W invd, uint32_t flags) {
    typedef typename G::W W;
    const int rows = G::rows_in_lane(j);
    if (rows == 0) return;
    const int base = board_bit0 + j * G::RPL * G::N;
    const W ones = (W(1) << (rows * G::N)) - 1;       // rows*N < word bits by construction
    stream_put(s_bits, base + 0 * G::NP, compact_rows<G>(black));
    stream_put(s_bits, base + 1 * G::NP, compact_rows<G>(white));
    stream_put(s_bits, base + 2 * G::NP, (flags & FLAG_TURN) ? ones : W(0));
    stream_put(s_bits, base + 3 * G::NP, compact_rows<G>(invd));
    stream_put(s_bits, base + 4 * G::NP, (flags & FLAG_PASS) ? ones : W(0));
    stream_put(s_bits, base + 5 * G::NP, (flags & FLAG_DONE) ? ones : W(0));
}

// Expand stream bits [head, head+count) to f32 / u8.  Stream bit i belongs to element base[i]; `base` is
// 16-byte aligned and the first `head` bits (< one vector) are padding, so that every full vector is one
// aligned 128-bit store; only the first and last vector of a range can be partial (scalar stores).
// f32: one stream nibble -> one float4 through a 16-entry shared-memory table (1 LDS.128, no ALU expansion).
// With a stride of NT = 32*k threads a thread always reads the same nibble position of successive words, so
// the inner loop is: LDS word, shift (loop-invariant amount), mask, LDS.128 table, STG.128, two pointer bumps.
template <int NT>
__device__ __forceinline__ void emit_f32(const uint32_t* s_bits, const float4* s_lut, int head, int count, float* base,
                                         int tid) {
    static_assert(NT % 32 == 0, "stride must keep the nibble position fixed per thread");
    const int end = head + count;
    const int q_lo = (head + 3) >> 2, q_hi = end >> 2;             // full vectors are [q_lo, q_hi)
    float4* base4 = reinterpret_cast<float4*>(base);
    const int sh = (tid & 7) << 2;
    int q = tid;
    if (q < q_lo) q += NT;                                          // q_lo is 0 or 1
    const uint32_t* w = s_bits + (q >> 3);
    float4* g = base4 + q;
#pragma unroll 4
    for (; q < q_hi; q += NT, w += NT / 8, g += NT) __stcs(g, s_lut[(*w >> sh) & 15u]);
    if (tid < 8) {                                                  // the (at most two) partial vectors
        const int e = tid < 4 ? tid : (q_hi << 2) + tid - 4;        // elements 0..3 and the last vector's
        if (e >= head && e < end && (e < (q_lo << 2) || e >= (q_hi << 2)))
            base[e] = ((s_bits[e >> 5] >> (e & 31)) & 1u) ? 1.0f : 0.0f;
    }
}
// u8: 16 elements per 16-byte vector; each of the two stream bytes -> 8 output bytes through a 256-entry table.
template <int NT>
__device__ __forceinline__ void emit_u8(const uint32_t* s_bits, const uint2* s_lut, int head, int count, uint8_t* base, int tid) {
    static_assert(NT % 32 == 0, "stride must keep the half-word position fixed per thread");
    const int end = head + count;
    const int q_lo = (head + 15) >> 4, q_hi = end >> 4;             // full vectors are [q_lo, q_hi)
    uint4* base4 = reinterpret_cast<uint4*>(base);
    const int sh = (tid & 1) << 4;
    int q = tid;
    if (q < q_lo) q += NT;
    const uint32_t* w = s_bits + (q >> 1);
    uint4* g = base4 + q;
#pragma unroll 4
    for (; q < q_hi; q += NT, w += NT / 2, g += NT) {
        const uint32_t h = *w >> sh;
        const uint2 lo = s_lut[h & 255u], hi = s_lut[(h >> 8) & 255u];
        __stcs(g, make_uint4(lo.x, lo.y, hi.x, hi.y));
    }
    if (tid < 32) {                                                  // the (at most two) partial vectors
        const int e = tid < 16 ? tid : (q_hi << 4) + tid - 16;
        if (e >= head && e < end && (e < (q_lo << 4) || e >= (q_hi << 4)))
            base[e] = uint8_t((s_bits[e >> 5] >> (e & 31)) & 1u);
    }
}

// 16-bit floats (bf16 / f16): 8 elements per 16-byte vector; one stream byte -> two table lookups (4 halves each).
template <int NT>
__device__ __forceinline__ void emit_h16(const uint32_t* s_bits, const uint2* s_lut, int head, int count, uint16_t* base,
                                         int tid, uint16_t one) {
    static_assert(NT % 32 == 0, "stride must keep the byte position fixed per thread");
    const int end = head + count;
    const int q_lo = (head + 7) >> 3, q_hi = end >> 3;             // full vectors are [q_lo, q_hi)
    uint4* base4 = reinterpret_cast<uint4*>(base);
    const int sh = (tid & 3) << 3;
    int q = tid;
    if (q < q_lo) q += NT;
    const uint32_t* w = s_bits + (q >> 2);
    uint4* g = base4 + q;
#pragma unroll 4
    for (; q < q_hi; q += NT, w += NT / 4, g += NT) {
        const uint32_t byte = (*w >> sh) & 255u;
        const uint2 lo = s_lut[byte & 15u], hi = s_lut[byte >> 4];
        __stcs(g, make_uint4(lo.x, lo.y, hi.x, hi.y));
    }
    if (tid < 16) {                                                 // the (at most two) partial vectors
        const int e = tid < 8 ? tid : (q_hi << 3) + tid - 8;
        if (e >= head && e < end && (e < (q_lo << 3) || e >= (q_hi << 3)))
            base[e] = ((s_bits[e >> 5] >> (e & 31)) & 1u) ? one : uint16_t(0);
    }
}

// dtype dispatch shared by every kernel that writes observations.  One shared-memory table serves all dtypes:
// f32: 16 x float4 (nibble -> 4 floats); bf16/f16: 16 x uint2 (nibble -> 4 halves); u8: 256 x uint2 (byte -> 8 bytes).
constexpr int LUT_F4 = 128;                                          // 2 KB
__device__ __forceinline__ uint16_t obs_one16(int dt) { return dt == DT_BF16 ? uint16_t(0x3F80) : uint16_t(0x3C00); }
__device__ __forceinline__ int obs_align_mask(int dt) { return dt == DT_F32 ? 3 : (dt == DT_U8 ? 15 : 7); }
template <int NT>
__device__ __forceinline__ void obs_lut_init(float4* s_lut, int dt, int tid) {
    if (dt == DT_U8) {
        for (int i = tid; i < 256; i += NT)
            reinterpret_cast<uint2*>(s_lut)[i] = make_uint2(((i & 15u) * 0x00204081u) & 0x01010101u,   // 4 bits -> 4 bytes
                                                             (((i >> 4) & 15u) * 0x00204081u) & 0x01010101u);
        return;
    }
    if (tid >= 16) return;
    if (dt == DT_BF16 || dt == DT_F16) {
        const uint32_t one = obs_one16(dt);
        reinterpret_cast<uint2*>(s_lut)[tid] =
            make_uint2(((tid & 1) ? one : 0u) | ((tid & 2) ? one << 16 : 0u), ((tid & 4) ? one : 0u) | ((tid & 8) ? one << 16 : 0u));
    } else {
        s_lut[tid] = make_float4(float(tid & 1), float((tid >> 1) & 1), float((tid >> 2) & 1), float((tid >> 3) & 1));
    }
}
// expand stream bits [head, head+count) into the dense buffer `buf` starting at element index `at` (16-byte aligned)
template <int NT>
__device__ __forceinline__ void emit_obs(int dt, const uint32_t* s_bits, const float4* s_lut, int head, int count, void* buf,
                                         long long at, int tid) {
    if (dt == DT_F32) emit_f32<NT>(s_bits, s_lut, head, count, static_cast<float*>(buf) + at, tid);
    else if (dt == DT_U8) emit_u8<NT>(s_bits, reinterpret_cast<const uint2*>(s_lut), head, count, static_cast<uint8_t*>(buf) + at, tid);
    else emit_h16<NT>(s_bits, reinterpret_cast<const uint2*>(s_lut), head, count, static_cast<uint16_t*>(buf) + at, tid,
                      obs_one16(dt));
}

// =================================================================================================
// The hot kernel: one ply for a tile of boards (STEP), fused reset+sample+ply (ROLLOUT), or one child
// per (parent, action) slot (CHILDREN).
// =================================================================================================
template <class G, int MODE>
__global__ void __launch_bounds__(Tile<G>::THREADS) k_step(const StepArgs a) {
    typedef typename G::W W;
    typedef Tile<G> T;
    __shared__ __align__(16) uint32_t s_rec[T::BT * G::REC_W32];
    __shared__ uint32_t s_bits[T::STREAM_W32];
    __shared__ __align__(16) float4 s_lut[LUT_F4];
    __shared__ __align__(8) uint64_t s_bar;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const long long tile_base = (long long)blockIdx.x * T::BT;
    const long long left = a.slots - tile_base;
    const int nb = left < T::BT ? int(left) : T::BT;             // slots of this tile that exist
    const bool want_obs = a.obs != nullptr;

    if (MODE != MODE_CHILDREN) {
        if (tid == 0) {
            mbar_init(&s_bar, 1);
            fence_mbar_init();
        }
    }
    if (want_obs) {
        for (int i = tid; i < T::STREAM_W32; i += T::THREADS) s_bits[i] = 0;
        obs_lut_init<T::THREADS>(s_lut, a.obs_dtype, tid);
    }
    __syncthreads();
    if (MODE != MODE_CHILDREN) {
        if (tid == 0) {
            mbar_expect_tx(&s_bar, uint32_t(nb) * G::REC_BYTES);
            bulk_g2s(s_rec, a.rec_in + tile_base * G::REC_W32, uint32_t(nb) * G::REC_BYTES, &s_bar);
        }
        mbar_wait(&s_bar, 0);
    }

    const int slot_in_warp = lane / G::LPB;
    const int slot_local = warp * G::BPW + slot_in_warp;
    const bool real = slot_in_warp < G::BPW && slot_local < nb;
    const long long slot = tile_base + slot_local;
    DevOps<G> o;
    o.init(lane, real);
    const int j = o.j;
    const bool holder = real && G::rows_in_lane(j) > 0;          // lane owns words of the record
    uint32_t* my_rec = s_rec + slot_local * G::REC_W32;

    W black = 0, white = 0, invd = 0;
    uint32_t flags = 0;
    int action = G::NP;
    long long parent = 0;
    bool restarted = false;
    if (MODE == MODE_CHILDREN) {
        if (real) {
            parent = slot / G::A;
            action = int(slot - parent * G::A);
            const uint32_t* prec = a.rec_in + parent * G::REC_W32;
            if (holder) {
                black = rec_word<G>(prec, 0, j);
                white = rec_word<G>(prec, 1, j);
                invd = rec_word<G>(prec, 2, j);
            }
            flags = prec[G::FLAGS_IDX];
        }
    } else if (real) {
        if (holder) {
            black = rec_word<G>(my_rec, 0, j);
            white = rec_word<G>(my_rec, 1, j);
            invd = rec_word<G>(my_rec, 2, j);
        }
        flags = my_rec[G::FLAGS_IDX];
        if (MODE == MODE_STEP) {
            action = a.actions_in[slot];
            if ((a.opts & OPT_AUTO_RESET) && (flags & FLAG_DONE)) {   // vector-env style: a finished board restarts
                black = white = invd = 0;                              // from gogame.init_state before its action
                flags = 0;
                restarted = true;
                if (a.opts & OPT_RESET_SKIPS_ACTION) action = -1;      // gymnasium next-step autoreset: the action of
            }                                                          // the reset step is ignored (refused below)
        }
    }

    uint32_t opts = a.opts;
    bool child_valid = true;
    if (MODE == MODE_ROLLOUT) {
        if (flags & FLAG_DONE) {                                  // auto-reset (gogame.init_state)
            black = white = invd = 0;
            flags = 0;
        }
        const unsigned long long gb = a.board0 + (unsigned long long)slot;
        const uint32_t rnd = philox4x32_10(uint32_t(gb), uint32_t(gb >> 32), uint32_t(a.t), uint32_t(a.t >> 32),
                                           uint32_t(a.seed), uint32_t(a.seed >> 32));
        action = Algo<DevOps<G>>::sample_action(o, G(), invd, rnd);
        opts = 0;
    }
    if (MODE == MODE_CHILDREN) {
        // gogame.valid_moves: everything is "valid" once the game has ended (gogame.py:155-156)
        const bool hit = o.any_board(o.single(action < G::NP ? action : 0) & invd);   // collective: no short-circuit
        const bool on_invd = action < G::NP && hit;
        child_valid = real && ((flags & FLAG_DONE) || !on_invd);
        if (!child_valid) action = -1;                            // refused below -> planes untouched
        opts &= OPT_CANONICAL;
    }

    int status = Algo<DevOps<G>>::step(o, G(), black, white, invd, flags, action, opts);
    if (MODE == MODE_STEP && restarted && (a.opts & OPT_RESET_SKIPS_ACTION)) status = ST_OK;   // fresh board, nothing played

    if (MODE == MODE_CHILDREN) {
        if (status != ST_OK) {                                    // padded slot: all zeros
            black = white = invd = 0;
            flags = 0;
            if (child_valid && a.status && j == 0) a.status[parent] = 1;   // reference would assert here
        }
        if (real && j == 0 && a.valid_out) a.valid_out[slot] = child_valid ? 1 : 0;
    }
    if (real && j == 0) {
        if (MODE == MODE_STEP && a.status) a.status[slot] = uint8_t(status);
        if (MODE == MODE_ROLLOUT && a.actions_out) a.actions_out[slot] = action;
        if (a.done_out) a.done_out[slot] = (flags & FLAG_DONE) ? 1 : 0;
    }
    // Scoring epilogue (gogame.areas + GoEnv.reward).  REAL rewards need areas only on finished boards, so the
    // two floods are skipped unless some board of this warp just ended (warp-uniform vote).
    const bool over = (flags & FLAG_DONE) != 0;
    const bool want_reward = a.reward_out != nullptr;
    const bool need_areas = a.areas_out != nullptr || (want_reward && (a.reward_mode == 2 || __any_sync(0xffffffffu, over)));
    if (need_areas) {
        int ba, wa;
        Algo<DevOps<G>>::areas(o, black, white, ba, wa);
        if (real && j == 0) {
            if (a.areas_out) {
                a.areas_out[2 * slot] = ba;
                a.areas_out[2 * slot + 1] = wa;
            }
            if (want_reward) {
                const float diff = float(ba - wa) - a.komi;
                float r;
                if (a.reward_mode == 1) r = over ? (diff > 0.f ? 1.f : (diff < 0.f ? -1.f : 0.f)) : 0.f;
                else r = over ? (diff > 0.f ? float(G::NP) : -float(G::NP)) : diff;
                a.reward_out[slot] = r;
            }
        }
    } else if (want_reward && real && j == 0) {
        a.reward_out[slot] = 0.f;
    }

    // new record -> shared tile (padding words: kept from the input for STEP/ROLLOUT, zeroed for CHILDREN).
    // Every lane of a board read the flags word above; order those reads before lane 0 overwrites it.
    __syncwarp();
    if (holder) {
        rec_word_store<G>(my_rec, 0, j, black);
        rec_word_store<G>(my_rec, 1, j, white);
        rec_word_store<G>(my_rec, 2, j, invd);
    }
    if (real && j == 0) {
        my_rec[G::FLAGS_IDX] = flags;
        if (MODE == MODE_CHILDREN)
            for (int i = G::FLAGS_IDX + 1; i < G::REC_W32; ++i) my_rec[i] = 0;
    }
    if (want_obs && holder) stream_put_board<G>(s_bits, slot_local * T::DENSE, j, black, white, invd, flags);
    fence_proxy_async();            // generic-proxy writes to s_rec must be visible to the bulk store
    __syncthreads();

    if (a.rec_out && tid == 0) {
        bulk_s2g(a.rec_out + tile_base * G::REC_W32, s_rec, uint32_t(nb) * G::REC_BYTES);
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
    if (want_obs) {
        const int elems = nb * T::DENSE;
        const long long ebase = tile_base * T::DENSE;
        emit_obs<T::THREADS>(a.obs_dtype, s_bits, s_lut, 0, elems, a.obs, ebase, tid);
    }
    if (a.rec_out && tid == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

// =================================================================================================
// Persistent rollout kernel: `plies` consecutive fused plies per launch.  The boards of a warp stay in
// REGISTERS for the whole launch (records are staged in by one TMA bulk load per CTA and leave by one bulk
// store at the end); every ply each warp resets its finished boards, samples, plays, and expands the dense
// observation of ITS boards from a warp-private bit stream - no CTA barrier inside the ply loop, so warps
// drift apart and the HBM-bound observation stores of some warps overlap the ALU-bound rules of others.
// =================================================================================================
struct RolloutArgs {
    uint32_t* rec;              // [B] records, updated in place
    long long boards;
    unsigned long long seed, board0, t0;
    int plies;
    int32_t* actions_log;       // [plies, B] or NULL
    uint8_t* done_log;          // [plies, B] or NULL
    float* reward_log;          // [plies, B] or NULL
    int reward_mode;
    float komi;
    void* obs_ring;             // [ring, B, 6, N, N] or NULL; ply t writes slot t % ring
    int obs_dtype, ring;
    int variant;                // 0: k_rollout (lane-sliced boards), 1: k_rollout_tpb (thread per board)
    // Dynamic scheduling (ws != NULL): the launch is cut into `rounds` blocks of `block_plies` plies; the grid has
    // tiles * rounds CTAs, each takes a ticket (atomicAdd on ws[0]) = (round, tile) in round-major order, waits until
    // the tile's previous block is finished (ws[1 + tile] >= round), plays its block and publishes ws[1 + tile] =
    // round + 1.  Tickets are handed out in start order, so the CTA a ticket waits for is always running or done.
    // The hardware block scheduler thereby balances the (persistently unequal) per-tile work over the SMs.
    int* ws;
    int block_plies, rounds;
    long long tiles;
};

// (round, tile, ply range) of this CTA; static scheduling: the whole launch for tile blockIdx.x
struct WorkItem {
    long long tile;
    int round, p_lo, p_hi;
};
__device__ __forceinline__ WorkItem take_work(const RolloutArgs& a, int* s_ticket, int tid) {
    WorkItem w;
    w.tile = blockIdx.x;
    w.round = 0;
    w.p_lo = 0;
    w.p_hi = a.plies;
    if (a.ws) {
        if (tid == 0) *s_ticket = atomicAdd(a.ws, 1);
        __syncthreads();
        const long long ticket = *s_ticket;
        w.round = int(ticket / a.tiles);
        w.tile = ticket - (long long)w.round * a.tiles;
        w.p_lo = w.round * a.block_plies;
        w.p_hi = w.p_lo + a.block_plies < a.plies ? w.p_lo + a.block_plies : a.plies;
        if (w.round > 0 && tid == 0) {
            while (ld_acquire_gpu(a.ws + 1 + w.tile) < w.round) __nanosleep(200);
            fence_proxy_async_all();                               // the records arrive through the async proxy (TMA)
        }
    }
    return w;
}
__device__ __forceinline__ void publish_work(const RolloutArgs& a, const WorkItem& w) {   // one thread, after its bulk
    if (a.ws) {                                                                           // store has completed
        fence_proxy_async_all();
        __threadfence();
        st_release_gpu(a.ws + 1 + w.tile, w.round + 1);
    }
}

template <class G>
struct WarpStream {
    static constexpr int DENSE = 6 * G::NP;
    static constexpr int W32 = (G::BPW * DENSE + 15 + 31) / 32 + 2;     // + up to 15 padding bits in front
};

template <class G>
__global__ void __launch_bounds__(Tile<G>::THREADS, Tile<G>::ROLLOUT_MIN_BLOCKS) k_rollout(const RolloutArgs a) {
    typedef typename G::W W;
    typedef Tile<G> T;
    typedef WarpStream<G> WS;
    typedef DevOps<G> O;
    __shared__ __align__(16) uint32_t s_rec[T::BT * G::REC_W32];
    __shared__ uint32_t s_bits_all[T::WPC][WS::W32];
    __shared__ __align__(16) float4 s_lut[LUT_F4];
    __shared__ __align__(8) uint64_t s_bar;
    __shared__ int s_ticket;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const WorkItem work = take_work(a, &s_ticket, tid);
    const long long tile_base = work.tile * T::BT;
    const long long left = a.boards - tile_base;
    const int nb = left < T::BT ? int(left) : T::BT;
    const bool want_obs = a.obs_ring != nullptr;

    if (tid == 0) {
        mbar_init(&s_bar, 1);
        fence_mbar_init();
    }
    obs_lut_init<T::THREADS>(s_lut, a.obs_dtype, tid);
    __syncthreads();                                               // also: thread 0 has seen the tile's previous block
    if (tid == 0) {
        mbar_expect_tx(&s_bar, uint32_t(nb) * G::REC_BYTES);
        bulk_g2s(s_rec, a.rec + tile_base * G::REC_W32, uint32_t(nb) * G::REC_BYTES, &s_bar);
    }
    mbar_wait(&s_bar, 0);

    const int slot_in_warp = lane / G::LPB;
    const int slot_local = warp * G::BPW + slot_in_warp;
    const bool real = slot_in_warp < G::BPW && slot_local < nb;
    const long long slot = tile_base + slot_local;
    O o;
    o.init(lane, real);
    const int j = o.j;
    const bool holder = real && G::rows_in_lane(j) > 0;
    uint32_t* my_rec = s_rec + slot_local * G::REC_W32;

    W black = 0, white = 0, invd = 0;
    uint32_t flags = 0;
    if (holder) {
        black = rec_word<G>(my_rec, 0, j);
        white = rec_word<G>(my_rec, 1, j);
        invd = rec_word<G>(my_rec, 2, j);
    }
    if (real) flags = my_rec[G::FLAGS_IDX];

    // this warp's slice of the dense output: boards [wb0, wb0 + nbw)
    const long long wb0 = tile_base + warp * G::BPW;
    int nbw = nb - warp * G::BPW;
    nbw = nbw < 0 ? 0 : (nbw > G::BPW ? G::BPW : nbw);
    const long long e0 = wb0 * WS::DENSE;                          // first element inside an observation slot
    const int align_mask = obs_align_mask(a.obs_dtype);            // elements per 16-byte vector, minus 1
    const int count = nbw * WS::DENSE;
    const long long slot_elems = a.boards * WS::DENSE;
    const unsigned long long gb = a.board0 + (unsigned long long)slot;

    for (int p = work.p_lo; p < work.p_hi; ++p) {
        const unsigned long long t = a.t0 + (unsigned long long)p;
        if (flags & FLAG_DONE) {                                   // auto-reset (gogame.init_state)
            black = white = invd = 0;
            flags = 0;
        }
        const uint32_t rnd = philox4x32_10(uint32_t(gb), uint32_t(gb >> 32), uint32_t(t), uint32_t(t >> 32),
                                           uint32_t(a.seed), uint32_t(a.seed >> 32));
        const int action = Algo<O>::sample_action(o, G(), invd, rnd);
        Algo<O>::step(o, G(), black, white, invd, flags, action, 0u);

        const bool over = (flags & FLAG_DONE) != 0;
        const long long log_at = (long long)p * a.boards + slot;
        if (real && j == 0) {
            if (a.actions_log) a.actions_log[log_at] = action;
            if (a.done_log) a.done_log[log_at] = over ? 1 : 0;
        }
        if (a.reward_log) {
            const bool need_areas = a.reward_mode == 2 || __any_sync(0xffffffffu, over);
            float r = 0.f;
            if (need_areas) {
                int ba, wa;
                Algo<O>::areas(o, black, white, ba, wa);
                const float diff = float(ba - wa) - a.komi;
                if (a.reward_mode == 1) r = over ? (diff > 0.f ? 1.f : (diff < 0.f ? -1.f : 0.f)) : 0.f;
                else r = over ? (diff > 0.f ? float(G::NP) : -float(G::NP)) : diff;
            }
            if (real && j == 0) a.reward_log[log_at] = r;
        }
        if (want_obs) {
            // absolute element index of this warp's first value; the slot stride need not be a multiple of 16 bytes
            const long long abs0 = (long long)(t % (unsigned long long)a.ring) * slot_elems + e0;
            const int head = int(abs0 & align_mask);
            const long long at = abs0 - head;
            uint32_t* s_bits = s_bits_all[warp];
            for (int i = lane; i < WS::W32; i += 32) s_bits[i] = 0;
            __syncwarp();
            if (holder) stream_put_board<G>(s_bits, head + slot_in_warp * WS::DENSE, j, black, white, invd, flags);
            __syncwarp();
            emit_obs<32>(a.obs_dtype, s_bits, s_lut, head, count, a.obs_ring, at, lane);
            __syncwarp();
        }
    }

    __syncwarp();                                                  // flags word: reads (all lanes) before the write
    if (holder) {
        rec_word_store<G>(my_rec, 0, j, black);
        rec_word_store<G>(my_rec, 1, j, white);
        rec_word_store<G>(my_rec, 2, j, invd);
    }
    if (real && j == 0) my_rec[G::FLAGS_IDX] = flags;
    fence_proxy_async();
    __syncthreads();
    if (tid == 0) {
        bulk_s2g(a.rec + tile_base * G::REC_W32, s_rec, uint32_t(nb) * G::REC_BYTES);
        bulk_commit_wait_all();
        publish_work(a, work);
    }
}

// =================================================================================================
// Thread-per-board variant of the persistent rollout kernel: the whole board lives in the registers of ONE thread (gg::ArrayOps), so the rules need no shuffles
// and no ballots and a warp carries 32 boards; loops are thread-local (SIMT masks finished boards off).
// Same staging (TMA bulk load/store of the record tile) and the same warp-private observation emission.
// =================================================================================================
template <class G>
struct TpbTile {
    // small boards: 2 warps per CTA (1,024 CTAs for 65,536 boards = 6.9 per SM, 91 registers);
    // uint64 boards: 1 warp per CTA with up to 255 registers per thread (19x19: 243, no spills)
    static constexpr int THREADS = G::WB == 32 ? 64 : 32;
    static constexpr int MIN_BLOCKS = G::WB == 32 ? 8 : 4;
    static constexpr int BT = THREADS;
    static constexpr int DENSE = 6 * G::NP;
    static constexpr int WSTREAM_W32 = (32 * DENSE + 15 + 31) / 32 + 2;
};

// Bit stream of a thread-per-board warp WITHOUT shared-memory atomics: every thread assembles the 6*N*N dense bits of
// its board in registers (all offsets inside a board are compile-time constants), shifts them to the board's position
// in the warp stream with funnel shifts, and stores whole words.  The one word a board shares with its successor is
// merged through a shuffle and stored by the lower lane, so no word is written twice and nothing needs zeroing.
template <class G>
struct TpbStream {
    static constexpr int DENSE = 6 * G::NP;
    static constexpr int KW = (DENSE + 31) / 32;                 // words of one board's bits before shifting
    // uint32 boards whose stream words hold bits of at most two boards; the rest keeps the atomicOr stream
    static constexpr bool SHUFFLED = DENSE >= 64 && G::WB == 32;
};
template <class G>
__device__ __forceinline__ void tpb_stream_put(uint32_t* s_bits, int off, int lane, bool real, const ArrayPlane<G>& black,
                                               const ArrayPlane<G>& white, const ArrayPlane<G>& invd, uint32_t flags) {
    typedef TpbStream<G> TS;
    constexpr int KW = TS::KW;
    uint32_t L[KW];
#pragma unroll
    for (int k = 0; k < KW; ++k) L[k] = 0;
#pragma unroll
    for (int ch = 0; ch < 6; ++ch) {
#pragma unroll
        for (int j = 0; j < G::LPB; ++j) {
            const int rows = G::N - j * G::RPL < G::RPL ? G::N - j * G::RPL : G::RPL;    // compile-time after unrolling
            if (rows <= 0) continue;
            const int nbits = rows * G::N, at = ch * G::NP + j * G::RPL * G::N;
            uint32_t v;
            if (ch == 0) v = uint32_t(compact_rows<G>(black.w[j]));
            else if (ch == 1) v = uint32_t(compact_rows<G>(white.w[j]));
            else if (ch == 3) v = uint32_t(compact_rows<G>(invd.w[j]));
            else v = (flags & (ch == 2 ? FLAG_TURN : (ch == 4 ? FLAG_PASS : FLAG_DONE))) ? ((1u << nbits) - 1u) : 0u;
            L[at >> 5] |= v << (at & 31);
            if ((at & 31) + nbits > 32) L[(at >> 5) + 1] |= v >> (32 - (at & 31));
        }
    }
    const int sh = off & 31, w0 = off >> 5;
    const int last = (sh + TS::DENSE - 1) >> 5;                    // index (from w0) of my last word: KW - 1 or KW
    const bool tail_partial = ((sh + TS::DENSE) & 31) != 0;        // ... which I share with the next board
    // word k of my shifted stream = funnel(L[k-1], L[k]); my first word's low `sh` bits belong to the previous board
    const uint32_t first = L[0] << sh;
    const uint32_t next_first = __shfl_down_sync(0xffffffffu, first, 1);      // lanes without a board carry zeros
    const uint32_t merge = (tail_partial && lane < 31) ? next_first : 0u;
    if (!real) return;
    uint32_t* dst = s_bits + w0;
    if (sh == 0 || lane == 0) dst[0] = first;                      // otherwise the previous lane stores this word
#pragma unroll
    for (int k = 1; k < KW - 1; ++k) dst[k] = __funnelshift_l(L[k - 1], L[k], sh);
    {
        const uint32_t wa = __funnelshift_l(L[KW - 2], L[KW - 1], sh);
        dst[KW - 1] = last == KW - 1 ? (wa | merge) : wa;
        if (last == KW) dst[KW] = __funnelshift_l(L[KW - 1], 0u, sh) | merge;
    }
}

template <class G>
__global__ void __launch_bounds__(TpbTile<G>::THREADS, TpbTile<G>::MIN_BLOCKS) k_rollout_tpb(const RolloutArgs a) {
    typedef typename G::W W;
    typedef TpbTile<G> T;
    typedef ArrayOps<G> O;
    typedef ArrayPlane<G> P;
    __shared__ __align__(16) uint32_t s_rec[T::BT * G::REC_W32];
    __shared__ uint32_t s_bits_all[T::THREADS / 32][T::WSTREAM_W32];
    __shared__ __align__(16) float4 s_lut[LUT_F4];
    __shared__ __align__(8) uint64_t s_bar;
    __shared__ int s_ticket;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const WorkItem work = take_work(a, &s_ticket, tid);
    const long long tile_base = work.tile * T::BT;
    const long long left = a.boards - tile_base;
    const int nb = left < T::BT ? int(left) : T::BT;
    const bool want_obs = a.obs_ring != nullptr;
    uint32_t* s_bits = s_bits_all[warp];

    if (tid == 0) {
        mbar_init(&s_bar, 1);
        fence_mbar_init();
    }
    obs_lut_init<T::THREADS>(s_lut, a.obs_dtype, tid);
    __syncthreads();
    if (tid == 0) {
        mbar_expect_tx(&s_bar, uint32_t(nb) * G::REC_BYTES);
        bulk_g2s(s_rec, a.rec + tile_base * G::REC_W32, uint32_t(nb) * G::REC_BYTES, &s_bar);
    }
    mbar_wait(&s_bar, 0);

    const bool real = tid < nb;
    const long long slot = tile_base + tid;
    const O o(real);
    uint32_t* my_rec = s_rec + tid * G::REC_W32;
    P black = o.zero(), white = o.zero(), invd = o.zero();
    uint32_t flags = 0;
    if (real) {
#pragma unroll
        for (int j = 0; j < G::LPB; ++j) {
            black.w[j] = rec_word<G>(my_rec, 0, j);
            white.w[j] = rec_word<G>(my_rec, 1, j);
            invd.w[j] = rec_word<G>(my_rec, 2, j);
        }
        flags = my_rec[G::FLAGS_IDX];
    }

    const long long wb0 = tile_base + warp * 32;
    int nbw = nb - warp * 32;
    nbw = nbw < 0 ? 0 : (nbw > 32 ? 32 : nbw);
    const long long e0 = wb0 * T::DENSE;
    const int align_mask = obs_align_mask(a.obs_dtype);
    const int count = nbw * T::DENSE;
    const long long slot_elems = a.boards * T::DENSE;
    const unsigned long long gb = a.board0 + (unsigned long long)slot;

    for (int p = work.p_lo; p < work.p_hi; ++p) {
        const unsigned long long t = a.t0 + (unsigned long long)p;
        if (flags & FLAG_DONE) {
            black = white = invd = o.zero();
            flags = 0;
        }
        const uint32_t rnd = philox4x32_10(uint32_t(gb), uint32_t(gb >> 32), uint32_t(t), uint32_t(t >> 32),
                                           uint32_t(a.seed), uint32_t(a.seed >> 32));
        const int action = Algo<O>::sample_action(o, G(), invd, rnd);
        Algo<O>::step(o, G(), black, white, invd, flags, action, 0u);

        const bool over = (flags & FLAG_DONE) != 0;
        const long long log_at = (long long)p * a.boards + slot;
        if (real) {
            if (a.actions_log) a.actions_log[log_at] = action;
            if (a.done_log) a.done_log[log_at] = over ? 1 : 0;
            if (a.reward_log) {
                float r = 0.f;
                if (a.reward_mode == 2 || over) {
                    int ba, wa;
                    Algo<O>::areas(o, black, white, ba, wa);
                    const float diff = float(ba - wa) - a.komi;
                    if (a.reward_mode == 1) r = over ? (diff > 0.f ? 1.f : (diff < 0.f ? -1.f : 0.f)) : 0.f;
                    else r = over ? (diff > 0.f ? float(G::NP) : -float(G::NP)) : diff;
                }
                a.reward_log[log_at] = r;
            }
        }
        if (want_obs) {
            const long long abs0 = (long long)(t % (unsigned long long)a.ring) * slot_elems + e0;
            const int head = int(abs0 & align_mask);
            const long long at = abs0 - head;
            __syncwarp();                                          // the previous ply's expansion has read the stream
            if constexpr (TpbStream<G>::SHUFFLED) {
                tpb_stream_put<G>(s_bits, head + lane * T::DENSE, lane, real, black, white, invd, flags);
            } else {
                for (int i = lane; i < T::WSTREAM_W32; i += 32) s_bits[i] = 0;
                __syncwarp();
                if (real) {
#pragma unroll
                    for (int j = 0; j < G::LPB; ++j)
                        stream_put_board<G>(s_bits, head + lane * T::DENSE, j, black.w[j], white.w[j], invd.w[j], flags);
                }
            }
            __syncwarp();
            emit_obs<32>(a.obs_dtype, s_bits, s_lut, head, count, a.obs_ring, at, lane);
        }
    }

    __syncwarp();
    if (real) {
#pragma unroll
        for (int j = 0; j < G::LPB; ++j) {
            rec_word_store<G>(my_rec, 0, j, black.w[j]);
            rec_word_store<G>(my_rec, 1, j, white.w[j]);
            rec_word_store<G>(my_rec, 2, j, invd.w[j]);
        }
        my_rec[G::FLAGS_IDX] = flags;
    }
    fence_proxy_async();
    __syncthreads();
    if (tid == 0) {
        bulk_s2g(a.rec + tile_base * G::REC_W32, s_rec, uint32_t(nb) * G::REC_BYTES);
        bulk_commit_wait_all();
        publish_work(a, work);
    }
}

// =================================================================================================
// Thread-per-board form of the single-ply STEP kernel (caller-provided actions): the host-driven path of small boards.
// Same contract as k_step<G, MODE_STEP> (options, status, done / areas / reward outputs, refused boards untouched) and
// bit-identical results; the rules run without shuffles or ballots (gg::ArrayOps, 153 instead of 206 warp-instructions
// per 9x9 board-ply) and every warp expands the observation of ITS 32 boards from a warp-private stream - no CTA
// barrier between the rules and the stores.
// =================================================================================================
template <class G>
__global__ void __launch_bounds__(TpbTile<G>::THREADS, TpbTile<G>::MIN_BLOCKS) k_step_tpb(const StepArgs a) {
    typedef TpbTile<G> T;
    typedef ArrayOps<G> O;
    typedef ArrayPlane<G> P;
    __shared__ __align__(16) uint32_t s_rec[T::BT * G::REC_W32];
    __shared__ uint32_t s_bits_all[T::THREADS / 32][T::WSTREAM_W32];
    __shared__ __align__(16) float4 s_lut[LUT_F4];
    __shared__ __align__(8) uint64_t s_bar;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const long long tile_base = (long long)blockIdx.x * T::BT;
    const long long left = a.slots - tile_base;
    const int nb = left < T::BT ? int(left) : T::BT;
    const bool want_obs = a.obs != nullptr;
    if (tid == 0) {
        mbar_init(&s_bar, 1);
        fence_mbar_init();
    }
    if (want_obs) obs_lut_init<T::THREADS>(s_lut, a.obs_dtype, tid);
    __syncthreads();
    if (tid == 0) {
        mbar_expect_tx(&s_bar, uint32_t(nb) * G::REC_BYTES);
        bulk_g2s(s_rec, a.rec_in + tile_base * G::REC_W32, uint32_t(nb) * G::REC_BYTES, &s_bar);
    }
    mbar_wait(&s_bar, 0);

    const bool real = tid < nb;
    const long long slot = tile_base + tid;
    const O o(real);
    uint32_t* my_rec = s_rec + tid * G::REC_W32;
    P black = o.zero(), white = o.zero(), invd = o.zero();
    uint32_t flags = 0;
    int action = G::NP;
    bool restarted = false;
    if (real) {
#pragma unroll
        for (int j = 0; j < G::LPB; ++j) {
            black.w[j] = rec_word<G>(my_rec, 0, j);
            white.w[j] = rec_word<G>(my_rec, 1, j);
            invd.w[j] = rec_word<G>(my_rec, 2, j);
        }
        flags = my_rec[G::FLAGS_IDX];
        action = a.actions_in[slot];
        if ((a.opts & OPT_AUTO_RESET) && (flags & FLAG_DONE)) {       // a finished board restarts before its action
            black = white = invd = o.zero();
            flags = 0;
            restarted = true;
            if (a.opts & OPT_RESET_SKIPS_ACTION) action = -1;          // gymnasium next-step autoreset (refused below)
        }
    }
    int status = Algo<O>::step(o, G(), black, white, invd, flags, action, a.opts);
    if (restarted && (a.opts & OPT_RESET_SKIPS_ACTION)) status = ST_OK;    // fresh board, nothing played

    const bool over = (flags & FLAG_DONE) != 0;
    if (real) {
        if (a.status) a.status[slot] = uint8_t(status);
        if (a.done_out) a.done_out[slot] = over ? 1 : 0;
        const bool want_reward = a.reward_out != nullptr;
        if (a.areas_out != nullptr || (want_reward && (a.reward_mode == 2 || over))) {
            int ba, wa;
            Algo<O>::areas(o, black, white, ba, wa);
            if (a.areas_out) {
                a.areas_out[2 * slot] = ba;
                a.areas_out[2 * slot + 1] = wa;
            }
            if (want_reward) {
                const float diff = float(ba - wa) - a.komi;
                float r;
                if (a.reward_mode == 1) r = over ? (diff > 0.f ? 1.f : (diff < 0.f ? -1.f : 0.f)) : 0.f;
                else r = over ? (diff > 0.f ? float(G::NP) : -float(G::NP)) : diff;
                a.reward_out[slot] = r;
            }
        } else if (want_reward) {
            a.reward_out[slot] = 0.f;
        }
#pragma unroll
        for (int j = 0; j < G::LPB; ++j) {                             // padding words of the record keep the input's
            rec_word_store<G>(my_rec, 0, j, black.w[j]);
            rec_word_store<G>(my_rec, 1, j, white.w[j]);
            rec_word_store<G>(my_rec, 2, j, invd.w[j]);
        }
        my_rec[G::FLAGS_IDX] = flags;
    }
    fence_proxy_async();            // generic-proxy writes to s_rec must be visible to the bulk store
    __syncthreads();
    if (a.rec_out && tid == 0) {
        bulk_s2g(a.rec_out + tile_base * G::REC_W32, s_rec, uint32_t(nb) * G::REC_BYTES);
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
    if (want_obs) {
        uint32_t* s_bits = s_bits_all[warp];
        int nbw = nb - warp * 32;
        nbw = nbw < 0 ? 0 : (nbw > 32 ? 32 : nbw);
        const long long abs0 = (tile_base + warp * 32) * T::DENSE;    // this warp's first element of the dense output
        const int head = int(abs0 & obs_align_mask(a.obs_dtype));
        if constexpr (TpbStream<G>::SHUFFLED) {
            tpb_stream_put<G>(s_bits, head + lane * T::DENSE, lane, real, black, white, invd, flags);
        } else {
            for (int i = lane; i < T::WSTREAM_W32; i += 32) s_bits[i] = 0;
            __syncwarp();
            if (real) {
#pragma unroll
                for (int j = 0; j < G::LPB; ++j)
                    stream_put_board<G>(s_bits, head + lane * T::DENSE, j, black.w[j], white.w[j], invd.w[j], flags);
            }
        }
        __syncwarp();
        emit_obs<32>(a.obs_dtype, s_bits, s_lut, head, nbw * T::DENSE, a.obs, abs0 - head, lane);
    }
    if (a.rec_out && tid == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

inline unsigned blocks_for(long long items, int per_block) { return unsigned((items + per_block - 1) / per_block); }

// fills in the dynamic-scheduling fields for tiles of `tile_boards` boards; -> CTAs to launch
inline unsigned schedule(RolloutArgs& a, int tile_boards) {
    a.tiles = (a.boards + tile_boards - 1) / tile_boards;
    a.rounds = 1;
    if (a.ws && a.block_plies > 0 && a.plies >= 2 * a.block_plies) a.rounds = (a.plies + a.block_plies - 1) / a.block_plies;
    else a.ws = nullptr;
    return unsigned(a.tiles * a.rounds);
}

template <class G>
struct LaunchTpb {
    static void go(RolloutArgs a, cudaStream_t s) {
        const unsigned grid = schedule(a, TpbTile<G>::BT);
        k_rollout_tpb<G><<<grid, TpbTile<G>::THREADS, 0, s>>>(a);
    }
};

// ------------------------------------------------------------------ warp-layout helper kernels
template <class G>
struct WarpSlot {
    static constexpr int THREADS = 128;
    static constexpr int BOARDS = (THREADS / 32) * G::BPW;
};

template <class G>
__global__ void __launch_bounds__(WarpSlot<G>::THREADS) k_areas(const uint32_t* rec, long long batch, int32_t* out) {
    typedef typename G::W W;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int slot_in_warp = lane / G::LPB;
    const long long b = (long long)blockIdx.x * WarpSlot<G>::BOARDS + warp * G::BPW + slot_in_warp;
    const bool real = slot_in_warp < G::BPW && b < batch;
    DevOps<G> o;
    o.init(lane, real);
    W black = 0, white = 0;
    if (real && G::rows_in_lane(o.j) > 0) {
        black = rec_word<G>(rec + b * G::REC_W32, 0, o.j);
        white = rec_word<G>(rec + b * G::REC_W32, 1, o.j);
    }
    int ba, wa;
    Algo<DevOps<G>>::areas(o, black, white, ba, wa);
    if (real && o.j == 0) {
        out[2 * b] = ba;
        out[2 * b + 1] = wa;
    }
}

template <class G>
__global__ void __launch_bounds__(WarpSlot<G>::THREADS)
    k_sample(const uint32_t* rec, long long batch, unsigned long long seed, unsigned long long board0,
             unsigned long long t, int32_t* actions) {
    typedef typename G::W W;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int slot_in_warp = lane / G::LPB;
    const long long b = (long long)blockIdx.x * WarpSlot<G>::BOARDS + warp * G::BPW + slot_in_warp;
    const bool real = slot_in_warp < G::BPW && b < batch;
    DevOps<G> o;
    o.init(lane, real);
    W invd = 0;
    if (real && G::rows_in_lane(o.j) > 0) invd = rec_word<G>(rec + b * G::REC_W32, 2, o.j);
    const unsigned long long gb = board0 + (unsigned long long)b;
    const uint32_t rnd = philox4x32_10(uint32_t(gb), uint32_t(gb >> 32), uint32_t(t), uint32_t(t >> 32), uint32_t(seed),
                                       uint32_t(seed >> 32));
    const int action = Algo<DevOps<G>>::sample_action(o, G(), invd, rnd);
    if (real && o.j == 0) actions[b] = action;
}

// state_utils.update_pieces (state_utils.py:159-180) as a stand-alone operation: groups of colour 1 - player[b] that
// touch a point of `touch` (a plane in record layout: the BLACK plane of touch[b]) and have no liberty are removed
// from rec[b]; the removed stones are returned as the BLACK plane of killed[b] (the rest of that record is zeroed).
template <class G>
__global__ void __launch_bounds__(WarpSlot<G>::THREADS)
    k_capture(uint32_t* rec, const uint32_t* touch, const int32_t* player, uint32_t* killed, long long batch) {
    typedef typename G::W W;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int slot_in_warp = lane / G::LPB;
    const long long b = (long long)blockIdx.x * WarpSlot<G>::BOARDS + warp * G::BPW + slot_in_warp;
    const bool real = slot_in_warp < G::BPW && b < batch;
    DevOps<G> o;
    o.init(lane, real);
    const bool holder = real && G::rows_in_lane(o.j) > 0;
    W black = 0, white = 0, tch = 0;
    bool white_moved = false;
    if (real) white_moved = player[b] != 0;
    if (holder) {
        black = rec_word<G>(rec + b * G::REC_W32, 0, o.j);
        white = rec_word<G>(rec + b * G::REC_W32, 1, o.j);
        tch = rec_word<G>(touch + b * G::REC_W32, 0, o.j) & o.full();
    }
    const W own = white_moved ? white : black, opp = white_moved ? black : white;
    const W dead = Algo<DevOps<G>>::captured(o, opp, own, tch);
    if (holder) {
        rec_word_store<G>(rec + b * G::REC_W32, white_moved ? 0 : 1, o.j, opp & ~dead);
        rec_word_store<G>(killed + b * G::REC_W32, 0, o.j, dead);
        rec_word_store<G>(killed + b * G::REC_W32, 1, o.j, W(0));
        rec_word_store<G>(killed + b * G::REC_W32, 2, o.j, W(0));
    }
    if (real && o.j == 0)
        for (int i = G::FLAGS_IDX; i < G::REC_W32; ++i) killed[b * G::REC_W32 + i] = 0;
}

// ------------------------------------------------------------------------------ codecs & misc
template <class G, class T>
__global__ void k_pack(const T* dense, long long batch, uint32_t* rec) {
    // one thread per (board, unit): units 0..3*LPB-1 = plane words, unit 3*LPB = flags + padding
    constexpr int UNITS = 3 * G::LPB + 1;
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= batch * UNITS) return;
    const long long b = gid / UNITS;
    const int u = int(gid - b * UNITS);
    const T* st = dense + b * 6 * G::NP;
    uint32_t* r = rec + b * G::REC_W32;
    if (u < 3 * G::LPB) {
        const int plane = u / G::LPB, j = u - plane * G::LPB;
        const int ch = plane == 2 ? 3 : plane;
        typename G::W w = 0;
        const int rows = G::rows_in_lane(j);
        for (int i = 0; i < rows; ++i)
            for (int c = 0; c < G::N; ++c)
                if (st[ch * G::NP + (j * G::RPL + i) * G::N + c] != T(0)) w |= typename G::W(1) << (i * G::S + c);
        rec_word_store<G>(r, plane, j, w);
    } else {
        bool turn = false, pass = false, done = true;
        for (int p = 0; p < G::NP; ++p) {
            turn |= st[2 * G::NP + p] != T(0);
            pass |= st[4 * G::NP + p] != T(0);
            done &= st[5 * G::NP + p] != T(0);
        }
        r[G::FLAGS_IDX] = (turn ? FLAG_TURN : 0u) | (pass ? FLAG_PASS : 0u) | (done ? FLAG_DONE : 0u);
        for (int i = G::FLAGS_IDX + 1; i < G::REC_W32; ++i) r[i] = 0;
    }
}

template <class G, class T>
__global__ void k_unpack(const uint32_t* rec, long long batch, T* dense) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= batch * 6 * G::NP) return;
    const long long b = e / (6 * G::NP);
    const int rem = int(e - b * 6 * G::NP);
    const int ch = rem / G::NP, pt = rem - ch * G::NP;
    const uint32_t* r = rec + b * G::REC_W32;
    bool v;
    if (ch == 2) v = r[G::FLAGS_IDX] & FLAG_TURN;
    else if (ch == 4) v = r[G::FLAGS_IDX] & FLAG_PASS;
    else if (ch == 5) v = r[G::FLAGS_IDX] & FLAG_DONE;
    else {
        const int plane = ch == 3 ? 2 : ch;
        const int row = pt / G::N, c = pt - row * G::N, j = row / G::RPL;
        v = (rec_word<G>(r, plane, j) >> ((row - j * G::RPL) * G::S + c)) & 1;
    }
    dense[e] = v ? T(1) : T(0);
}

template <class G>
__global__ void k_unpack16(const uint32_t* rec, long long batch, uint16_t one, uint16_t* dense) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= batch * 6 * G::NP) return;
    const long long b = e / (6 * G::NP);
    const int rem = int(e - b * 6 * G::NP);
    const int ch = rem / G::NP, pt = rem - ch * G::NP;
    const uint32_t* r = rec + b * G::REC_W32;
    bool v;
    if (ch == 2) v = r[G::FLAGS_IDX] & FLAG_TURN;
    else if (ch == 4) v = r[G::FLAGS_IDX] & FLAG_PASS;
    else if (ch == 5) v = r[G::FLAGS_IDX] & FLAG_DONE;
    else {
        const int plane = ch == 3 ? 2 : ch;
        const int row = pt / G::N, c = pt - row * G::N, j = row / G::RPL;
        v = (rec_word<G>(r, plane, j) >> ((row - j * G::RPL) * G::S + c)) & 1;
    }
    dense[e] = v ? one : uint16_t(0);
}

template <class G, class T>
__global__ void k_valid(const uint32_t* rec, long long batch, int ended_quirk, T* out) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= batch * G::A) return;
    const long long b = e / G::A;
    const int act = int(e - b * G::A);
    const uint32_t* r = rec + b * G::REC_W32;
    bool ok = true;
    if (act < G::NP && !(ended_quirk && (r[G::FLAGS_IDX] & FLAG_DONE))) {
        const int row = act / G::N, c = act - row * G::N, j = row / G::RPL;
        ok = !((rec_word<G>(r, 2, j) >> ((row - j * G::RPL) * G::S + c)) & 1);
    }
    out[e] = ok ? T(1) : T(0);
}

template <class G>
__global__ void k_reset(uint32_t* rec, long long batch, const uint8_t* mask) {
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= batch * G::REC_W32) return;
    const long long b = gid / G::REC_W32;
    if (mask == nullptr || mask[b]) rec[gid] = 0;
}

template <class G>
__global__ void k_canonical(const uint32_t* in, uint32_t* out, long long batch) {
    // one thread per board so that in-place use (in == out) is race-free
    const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= batch) return;
    const uint32_t* r = in + b * G::REC_W32;
    uint32_t* w = out + b * G::REC_W32;
    const uint32_t flags = r[G::FLAGS_IDX];
    const bool swap = flags & FLAG_TURN;
    for (int i = 0; i < G::PLANE_W32; ++i) {
        const uint32_t bl = r[i], wh = r[i + G::PLANE_W32];
        w[i] = swap ? wh : bl;
        w[i + G::PLANE_W32] = swap ? bl : wh;
    }
    for (int i = 2 * G::PLANE_W32; i < G::REC_W32; ++i) w[i] = r[i];
    w[G::FLAGS_IDX] = flags & ~uint32_t(FLAG_TURN);
}

// One of the 8 dihedral transforms on packed records, numbered like gogame.all_symmetries (gogame.py:358-382):
// bit 0 of `sym` mirrors the columns, bit 1 mirrors the rows, bit 2 turns the result a quarter counter-clockwise
// (np.rot90), applied in that order.  One thread per (board, plane word), flags copied.
template <class G>
__global__ void k_symmetry(const uint32_t* in, uint32_t* out, long long batch, int sym) {
    constexpr int UNITS = 3 * G::LPB + 1;
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= batch * UNITS) return;
    const long long b = gid / UNITS;
    const int u = int(gid - b * UNITS);
    const uint32_t* r = in + b * G::REC_W32;
    uint32_t* w = out + b * G::REC_W32;
    if (u == 3 * G::LPB) {
        for (int i = G::FLAGS_IDX; i < G::REC_W32; ++i) w[i] = r[i];
        return;
    }
    const int plane = u / G::LPB, j = u - plane * G::LPB;
    const bool mirror_cols = (sym & 1) != 0, mirror_rows = (sym & 2) != 0, quarter = (sym & 4) != 0;
    typename G::W word = 0;
    const int rows = G::rows_in_lane(j);
    for (int i = 0; i < rows; ++i) {
        const int row = j * G::RPL + i;
        for (int c = 0; c < G::N; ++c) {
            // output point (row, c) <- point (a, d) of the mirrored image <- point (sr, sc) of the input
            const int a = quarter ? c : row, d = quarter ? G::N - 1 - row : c;
            const int sr = mirror_rows ? G::N - 1 - a : a, sc = mirror_cols ? G::N - 1 - d : d;
            const int sj = sr / G::RPL;
            const typename G::W src = rec_word<G>(r, plane, sj);
            if ((src >> ((sr - sj * G::RPL) * G::S + sc)) & 1) word |= typename G::W(1) << (i * G::S + c);
        }
    }
    rec_word_store<G>(w, plane, j, word);
}

// ------------------------------------------------------------------------------ launch table
struct SizeVTable {
    int n, rec_bytes, lpb, rpl, wordbits, bpw, tile_boards, tile_threads;
    cudaError_t (*step)(const StepArgs&, int mode, cudaStream_t);
    cudaError_t (*rollout)(const RolloutArgs&, cudaStream_t);
    cudaError_t (*areas)(const uint32_t*, long long, int32_t*, cudaStream_t);
    cudaError_t (*sample)(const uint32_t*, long long, unsigned long long, unsigned long long, unsigned long long, int32_t*,
                          cudaStream_t);
    cudaError_t (*pack)(const void*, int dtype, long long, uint32_t*, cudaStream_t);
    cudaError_t (*unpack)(const uint32_t*, long long, int dtype, void*, cudaStream_t);
    cudaError_t (*valid)(const uint32_t*, long long, int quirk, int dtype, void*, cudaStream_t);
    cudaError_t (*reset)(uint32_t*, long long, const uint8_t*, cudaStream_t);
    cudaError_t (*canonical)(const uint32_t*, uint32_t*, long long, cudaStream_t);
    cudaError_t (*symmetry)(const uint32_t*, uint32_t*, long long, int, cudaStream_t);
    cudaError_t (*capture)(uint32_t*, const uint32_t*, const int32_t*, uint32_t*, long long, cudaStream_t);
};

template <class G>
struct Launch {
    // which STEP kernel: an explicit choice (GG_STEP_KERNEL_*), else thread-per-board for small boards (uint32 words,
    // <= 3 words per plane) in large batches (>= 48 Ki boards) unless a float32 observation is written.  Measured on
    // 9x9 x 65,536 (tools/kstep_phase_probe.py, us per launch, lanes -> thread): no observation 17.0 -> 15.2, u8
    // 20.6 -> 17.3, bf16 21.9 -> 18.9, f32 28.9 -> 29.3 (store-bound: the CTA-wide emission of k_step is as good).
    static bool step_uses_thread_kernel(const StepArgs& a) {
        if (a.opts & OPT_KERNEL_THREAD) return true;
        if (a.opts & OPT_KERNEL_LANES) return false;
        return G::WB == 32 && G::LPB <= 3 && a.slots >= 49152 && !(a.obs != nullptr && a.obs_dtype == DT_F32);
    }
    static cudaError_t step(const StepArgs& a, int mode, cudaStream_t s) {
        if (a.slots <= 0) return cudaSuccess;
        const unsigned grid = blocks_for(a.slots, Tile<G>::BT);
        if (mode == MODE_STEP && step_uses_thread_kernel(a)) {
            StepArgs b = a;
            b.opts &= ~(OPT_KERNEL_LANES | OPT_KERNEL_THREAD);
            k_step_tpb<G><<<blocks_for(a.slots, TpbTile<G>::BT), TpbTile<G>::THREADS, 0, s>>>(b);
        } else if (mode == MODE_STEP) k_step<G, MODE_STEP><<<grid, Tile<G>::THREADS, 0, s>>>(a);
        else if (mode == MODE_ROLLOUT) k_step<G, MODE_ROLLOUT><<<grid, Tile<G>::THREADS, 0, s>>>(a);
        else k_step<G, MODE_CHILDREN><<<grid, Tile<G>::THREADS, 0, s>>>(a);
        return cudaGetLastError();
    }
    static cudaError_t rollout(const RolloutArgs& a, cudaStream_t s) {
        if (a.boards <= 0 || a.plies <= 0) return cudaSuccess;
        if (a.variant == 1) {
            LaunchTpb<G>::go(a, s);
        } else {
            RolloutArgs b = a;
            const unsigned grid = schedule(b, Tile<G>::BT);
            k_rollout<G><<<grid, Tile<G>::THREADS, 0, s>>>(b);
        }
        return cudaGetLastError();
    }
    static cudaError_t capture(uint32_t* rec, const uint32_t* touch, const int32_t* player, uint32_t* killed, long long batch,
                               cudaStream_t s) {
        if (batch <= 0) return cudaSuccess;
        k_capture<G><<<blocks_for(batch, WarpSlot<G>::BOARDS), WarpSlot<G>::THREADS, 0, s>>>(rec, touch, player, killed, batch);
        return cudaGetLastError();
    }
    static cudaError_t areas(const uint32_t* rec, long long batch, int32_t* out, cudaStream_t s) {
        if (batch <= 0) return cudaSuccess;
        k_areas<G><<<blocks_for(batch, WarpSlot<G>::BOARDS), WarpSlot<G>::THREADS, 0, s>>>(rec, batch, out);
        return cudaGetLastError();
    }
    static cudaError_t sample(const uint32_t* rec, long long batch, unsigned long long seed, unsigned long long board0,
                              unsigned long long t, int32_t* actions, cudaStream_t s) {
        if (batch <= 0) return cudaSuccess;
        k_sample<G><<<blocks_for(batch, WarpSlot<G>::BOARDS), WarpSlot<G>::THREADS, 0, s>>>(rec, batch, seed, board0, t,
                                                                                          actions);
        return cudaGetLastError();
    }
    static cudaError_t pack(const void* dense, int dtype, long long batch, uint32_t* rec, cudaStream_t s) {
        if (batch <= 0) return cudaSuccess;
        const unsigned grid = blocks_for(batch * (3 * G::LPB + 1), 256);
        if (dtype == DT_U8) k_pack<G, uint8_t><<<grid, 256, 0, s>>>(static_cast<const uint8_t*>(dense), batch, rec);
        else if (dtype == DT_F32) k_pack<G, float><<<grid, 256, 0, s>>>(static_cast<const float*>(dense), batch, rec);
        else if (dtype == DT_BF16 || dtype == DT_F16)       // any non-zero bit pattern is a stone (0x8000 = -0 is not used)
            k_pack<G, uint16_t><<<grid, 256, 0, s>>>(static_cast<const uint16_t*>(dense), batch, rec);
        else k_pack<G, double><<<grid, 256, 0, s>>>(static_cast<const double*>(dense), batch, rec);
        return cudaGetLastError();
    }
    static cudaError_t unpack(const uint32_t* rec, long long batch, int dtype, void* dense, cudaStream_t s) {
        if (batch <= 0) return cudaSuccess;
        const unsigned grid = blocks_for(batch * 6 * G::NP, 256);
        if (dtype == DT_U8) k_unpack<G, uint8_t><<<grid, 256, 0, s>>>(rec, batch, static_cast<uint8_t*>(dense));
        else if (dtype == DT_F32) k_unpack<G, float><<<grid, 256, 0, s>>>(rec, batch, static_cast<float*>(dense));
        else if (dtype == DT_BF16 || dtype == DT_F16)
            k_unpack16<G><<<grid, 256, 0, s>>>(rec, batch, dtype == DT_BF16 ? uint16_t(0x3F80) : uint16_t(0x3C00),
                                               static_cast<uint16_t*>(dense));
        else k_unpack<G, double><<<grid, 256, 0, s>>>(rec, batch, static_cast<double*>(dense));
        return cudaGetLastError();
    }
    static cudaError_t valid(const uint32_t* rec, long long batch, int quirk, int dtype, void* out, cudaStream_t s) {
        if (batch <= 0) return cudaSuccess;
        const unsigned grid = blocks_for(batch * G::A, 256);
        if (dtype == DT_U8) k_valid<G, uint8_t><<<grid, 256, 0, s>>>(rec, batch, quirk, static_cast<uint8_t*>(out));
        else if (dtype == DT_F32) k_valid<G, float><<<grid, 256, 0, s>>>(rec, batch, quirk, static_cast<float*>(out));
        else k_valid<G, double><<<grid, 256, 0, s>>>(rec, batch, quirk, static_cast<double*>(out));
        return cudaGetLastError();
    }
    static cudaError_t reset(uint32_t* rec, long long batch, const uint8_t* mask, cudaStream_t s) {
        if (batch <= 0) return cudaSuccess;
        k_reset<G><<<blocks_for(batch * G::REC_W32, 256), 256, 0, s>>>(rec, batch, mask);
        return cudaGetLastError();
    }
    static cudaError_t canonical(const uint32_t* in, uint32_t* out, long long batch, cudaStream_t s) {
        if (batch <= 0) return cudaSuccess;
        k_canonical<G><<<blocks_for(batch, 128), 128, 0, s>>>(in, out, batch);
        return cudaGetLastError();
    }
    static cudaError_t symmetry(const uint32_t* in, uint32_t* out, long long batch, int sym, cudaStream_t s) {
        if (batch <= 0) return cudaSuccess;
        k_symmetry<G><<<blocks_for(batch * (3 * G::LPB + 1), 256), 256, 0, s>>>(in, out, batch, sym);
        return cudaGetLastError();
    }
    static constexpr SizeVTable table() {
        return SizeVTable{G::N, G::REC_BYTES, G::LPB, G::RPL, G::WB, G::BPW, Tile<G>::BT, Tile<G>::THREADS,
                          &step, &rollout, &areas, &sample, &pack, &unpack, &valid, &reset, &canonical, &symmetry, &capture};
    }
};

}  // namespace gg
