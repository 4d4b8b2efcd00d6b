// gg_sliced.cuh - generalised board decomposition: a board is spread over L lanes, K plane words per lane.
//   K = 1          -> DevOps   (one word per lane, gg_kernels.cuh)
//   K = LPB (L=1)  -> ArrayOps (whole board in one thread)
// In between (e.g. 19x19: 7 uint64 words as 4 lanes x 2 words, 8 boards per warp) the row neighbours of most
// words are registers of the same lane, so only one word in K needs a shuffle, and every plane operation offers
// K independent instructions to the scheduler.  Used by k_rollout_sliced (rollout variant 2).
#pragma once
#include "gg_kernels.cuh"

namespace gg {

template <class G, int K>
struct Slice {
    static constexpr int L = cdiv(G::LPB, K);          // lanes per board
    static constexpr int BPW = 32 / L;                 // boards per warp
    static constexpr unsigned GM = L >= 32 ? 0xffffffffu : ((1u << L) - 1u);
};

template <class G, int K>
struct SlicedPlane {
    typename G::W w[K];
};
template <class G, int K>
__device__ __forceinline__ SlicedPlane<G, K> operator|(SlicedPlane<G, K> a, const SlicedPlane<G, K>& b) {
#pragma unroll
    for (int i = 0; i < K; ++i) a.w[i] |= b.w[i];
    return a;
}
template <class G, int K>
__device__ __forceinline__ SlicedPlane<G, K> operator&(SlicedPlane<G, K> a, const SlicedPlane<G, K>& b) {
#pragma unroll
    for (int i = 0; i < K; ++i) a.w[i] &= b.w[i];
    return a;
}

template <class G, int K>
struct SlicedOps {
    typedef typename G::W W;
    typedef SlicedPlane<G, K> P;
    typedef Slice<G, K> S;
    static constexpr unsigned FULL = 0xffffffffu;
    int j;            // my lane inside the board's lane group
    unsigned gshift;  // first warp lane of the group
    P fullmask;

    __device__ __forceinline__ void init(int lane, bool real) {
        const int slot = lane / S::L;
        const bool ghost = slot >= S::BPW;
        gshift = ghost ? unsigned(S::BPW * S::L) : unsigned(slot * S::L);
        j = lane - int(gshift);
#pragma unroll
        for (int i = 0; i < K; ++i) {
            const int g = j * K + i;                   // global word index of my i-th word
            fullmask.w[i] = (!ghost && real && g < G::LPB) ? G::rows_mask(G::rows_in_lane(g)) : W(0);
        }
    }
    __device__ __forceinline__ P zero() const {
        P p;
#pragma unroll
        for (int i = 0; i < K; ++i) p.w[i] = 0;
        return p;
    }
    __device__ __forceinline__ P full() const { return fullmask; }
    __device__ __forceinline__ P andnot(P a, const P& b) const {
#pragma unroll
        for (int i = 0; i < K; ++i) a.w[i] &= ~b.w[i];
        return a;
    }
    __device__ __forceinline__ P east(P x) const {
#pragma unroll
        for (int i = 0; i < K; ++i) x.w[i] <<= 1;
        return x;
    }
    __device__ __forceinline__ P west(P x) const {
#pragma unroll
        for (int i = 0; i < K; ++i) x.w[i] >>= 1;
        return x;
    }
    __device__ __forceinline__ P south(const P& x) const {   // result[r] = x[r-1]
        P y;
        W below = 0;                                           // top row of the word before my first word
        if (S::L > 1) {
            const W prev = __shfl_up_sync(FULL, x.w[K - 1], 1);
            if (j > 0) below = prev >> ((G::RPL - 1) * G::S);
        }
#pragma unroll
        for (int i = 0; i < K; ++i) {
            const W in = G::RPL > 1 ? W(x.w[i] << (G::S % G::WB)) : W(0);
            y.w[i] = in | (i == 0 ? below : W(x.w[i ? i - 1 : 0] >> ((G::RPL - 1) * G::S)));
        }
        return y;
    }
    __device__ __forceinline__ P north(const P& x) const {   // result[r] = x[r+1]
        P y;
        W above = 0;                                           // bottom row of the word after my last word
        if (S::L > 1) {
            const W next = __shfl_down_sync(FULL, x.w[0], 1);
            if (j < S::L - 1) above = (next & G::row_bits()) << ((G::RPL - 1) * G::S);
        }
#pragma unroll
        for (int i = 0; i < K; ++i) {
            const W in = G::RPL > 1 ? W(x.w[i] >> (G::S % G::WB)) : W(0);
            y.w[i] = in | (i == K - 1 ? above : W((x.w[i + 1 < K ? i + 1 : i] & G::row_bits()) << ((G::RPL - 1) * G::S)));
        }
        return y;
    }
    __device__ __forceinline__ P rev(P x) const {
#pragma unroll
        for (int i = 0; i < K; ++i) x.w[i] = w_rev(x.w[i]);
        return x;
    }
    __device__ __forceinline__ P hfill(P s, const P& m, const P& mrev) const {
#pragma unroll
        for (int i = 0; i < K; ++i) s.w[i] = w_hfill(s.w[i], m.w[i], mrev.w[i]);
        return s;
    }
    __device__ __forceinline__ W fold(const P& x) const {
        W a = 0;
#pragma unroll
        for (int i = 0; i < K; ++i) a |= x.w[i];
        return a;
    }
    __device__ __forceinline__ bool any(const P& x) const { return __any_sync(FULL, fold(x) != 0); }
    __device__ __forceinline__ unsigned group_ballot(bool pred) const {
        return (__ballot_sync(FULL, pred) >> gshift) & S::GM;
    }
    __device__ __forceinline__ bool any_board(const P& x) const {
        if (S::L == 1) return fold(x) != 0;
        return group_ballot(fold(x) != 0) != 0;
    }
    __device__ __forceinline__ int lane_popc(const P& x) const {
        int c = 0;
#pragma unroll
        for (int i = 0; i < K; ++i) c += w_popc(x.w[i]);
        return c;
    }
    __device__ __forceinline__ int count2(const P& x) const {
        const int c = lane_popc(x);
        if (S::L == 1) return c > 2 ? 2 : c;
        const unsigned b1 = group_ballot(c >= 1), b2 = group_ballot(c >= 2);
        return (b2 || __popc(b1) >= 2) ? 2 : (b1 ? 1 : 0);
    }
    __device__ __forceinline__ int popc(const P& x) const {
        const int c = lane_popc(x);
        if (S::L == 1) return c;
        int total = 0;
#pragma unroll
        for (int l = 0; l < S::L; ++l) total += __shfl_sync(FULL, c, int(gshift) + l);
        return total;
    }
    __device__ __forceinline__ P lowest(const P& x) const {
        bool mine = true;
        if (S::L > 1) {
            const unsigned nz = group_ballot(fold(x) != 0);
            mine = j == __ffs(int(nz)) - 1;
        }
        P y;
        bool found = !mine;
#pragma unroll
        for (int i = 0; i < K; ++i) {
            y.w[i] = found ? W(0) : W(x.w[i] & (~x.w[i] + 1));
            found = found || x.w[i] != 0;
        }
        return y;
    }
    __device__ __forceinline__ P single(int pt) const {
        const int r = pt / G::N, c = pt - r * G::N;
        const int g = r / G::RPL;                                  // global word
        const W bit = W(1) << ((r - g * G::RPL) * G::S + c);
        P y;
#pragma unroll
        for (int i = 0; i < K; ++i) y.w[i] = (j * K + i == g) ? bit : W(0);
        return y;
    }
    __device__ __forceinline__ P pick(bool c, const P& a, const P& b) const {
        P y;
#pragma unroll
        for (int i = 0; i < K; ++i) y.w[i] = c ? a.w[i] : b.w[i];
        return y;
    }
    __device__ __forceinline__ int kth_point(const P& x, int k) const {
        const int c = lane_popc(x);
        int before = 0;
        if (S::L > 1) {
#pragma unroll
            for (int l = 0; l < S::L; ++l) {
                const int cl = __shfl_sync(FULL, c, int(gshift) + l);
                if (l < j) before += cl;
            }
        }
        int kl = k - before;
        const bool mine = kl >= 0 && kl < c;
        int pt = -1;
#pragma unroll
        for (int i = 0; i < K; ++i) {
            const int ci = w_popc(x.w[i]);
            if (mine && pt < 0 && kl >= 0 && kl < ci) {
                const int bit = w_select(x.w[i], kl);
                const int row = bit / G::S;
                pt = ((j * K + i) * G::RPL + row) * G::N + (bit - row * G::S);
            }
            kl -= ci;
        }
        if (S::L > 1) {
            const unsigned owner = group_ballot(mine);
            pt = __shfl_sync(FULL, pt, int(gshift) + (owner ? __ffs(int(owner)) - 1 : 0));
        }
        return pt;
    }
};

// ------------------------------------------------------------------------------------------------
// Persistent rollout kernel over SlicedOps<G, K> (same structure as k_rollout).
// ------------------------------------------------------------------------------------------------
template <class G, int K>
struct SlicedTile {
    typedef Slice<G, K> S;
    static constexpr int WARPS = 2;                              // small CTAs: 19x19 x 16,384 boards -> 1,024 CTAs
    static constexpr int THREADS = WARPS * 32;
    static constexpr int BT = WARPS * S::BPW;
    static constexpr int DENSE = 6 * G::NP;
    static constexpr int WSTREAM_W32 = (S::BPW * DENSE + 15 + 31) / 32 + 2;
    static constexpr int MIN_BLOCKS = 8;
};

template <class G, int K>
__global__ void __launch_bounds__(SlicedTile<G, K>::THREADS, SlicedTile<G, K>::MIN_BLOCKS)
    k_rollout_sliced(const RolloutArgs a) {
    typedef typename G::W W;
    typedef SlicedTile<G, K> T;
    typedef Slice<G, K> S;
    typedef SlicedOps<G, K> O;
    typedef SlicedPlane<G, K> P;
    __shared__ __align__(16) uint32_t s_rec[T::BT * G::REC_W32];
    __shared__ uint32_t s_bits_all[T::WARPS][T::WSTREAM_W32];
    __shared__ __align__(16) float4 s_lut[LUT_F4];
    __shared__ __align__(8) uint64_t s_bar;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const long long tile_base = (long long)blockIdx.x * T::BT;
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

    const int slot_in_warp = lane / S::L;
    const int slot_local = warp * S::BPW + slot_in_warp;
    const bool real = slot_in_warp < S::BPW && slot_local < nb;
    const long long slot = tile_base + slot_local;
    O o;
    o.init(lane, real);
    const int j = o.j;
    uint32_t* my_rec = s_rec + slot_local * G::REC_W32;

    P black = o.zero(), white = o.zero(), invd = o.zero();
    uint32_t flags = 0;
    if (real) {
#pragma unroll
        for (int i = 0; i < K; ++i) {
            const int g = j * K + i;
            if (g < G::LPB && G::rows_in_lane(g) > 0) {
                black.w[i] = rec_word<G>(my_rec, 0, g);
                white.w[i] = rec_word<G>(my_rec, 1, g);
                invd.w[i] = rec_word<G>(my_rec, 2, g);
            }
        }
        flags = my_rec[G::FLAGS_IDX];
    }

    const long long wb0 = tile_base + warp * S::BPW;
    int nbw = nb - warp * S::BPW;
    nbw = nbw < 0 ? 0 : (nbw > S::BPW ? S::BPW : nbw);
    const long long e0 = wb0 * T::DENSE;
    const int align_mask = obs_align_mask(a.obs_dtype);
    const int count = nbw * T::DENSE;
    const long long slot_elems = a.boards * T::DENSE;
    const unsigned long long gb = a.board0 + (unsigned long long)slot;

    for (int p = 0; p < a.plies; ++p) {
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
            const long long abs0 = (long long)(t % (unsigned long long)a.ring) * slot_elems + e0;
            const int head = int(abs0 & align_mask);
            const long long at = abs0 - head;
            __syncwarp();
            for (int i = lane; i < T::WSTREAM_W32; i += 32) s_bits[i] = 0;
            __syncwarp();
            if (real) {
#pragma unroll
                for (int i = 0; i < K; ++i) {
                    const int g = j * K + i;
                    if (g < G::LPB)
                        stream_put_board<G>(s_bits, head + slot_in_warp * T::DENSE, g, black.w[i], white.w[i], invd.w[i], flags);
                }
            }
            __syncwarp();
            emit_obs<32>(a.obs_dtype, s_bits, s_lut, head, count, a.obs_ring, at, lane);
            __syncwarp();
        }
    }

    __syncwarp();
    if (real) {
#pragma unroll
        for (int i = 0; i < K; ++i) {
            const int g = j * K + i;
            if (g < G::LPB && G::rows_in_lane(g) > 0) {
                rec_word_store<G>(my_rec, 0, g, black.w[i]);
                rec_word_store<G>(my_rec, 1, g, white.w[i]);
                rec_word_store<G>(my_rec, 2, g, invd.w[i]);
            }
        }
        if (j == 0) my_rec[G::FLAGS_IDX] = flags;
    }
    fence_proxy_async();
    __syncthreads();
    if (tid == 0) {
        bulk_s2g(a.rec + tile_base * G::REC_W32, s_rec, uint32_t(nb) * G::REC_BYTES);
        bulk_commit_wait_all();
    }
}

template <class G, int K, bool OK = (K > 1 && K < G::LPB)>
struct LaunchSliced {
    static bool go(const RolloutArgs& a, cudaStream_t s) {
        k_rollout_sliced<G, K><<<blocks_for(a.boards, SlicedTile<G, K>::BT), SlicedTile<G, K>::THREADS, 0, s>>>(a);
        return true;
    }
};
template <class G, int K>
struct LaunchSliced<G, K, false> {
    static bool go(const RolloutArgs&, cudaStream_t) { return false; }
};

}  // namespace gg
