// gg_array_ops.cuh - "whole board in one thread" Ops policy for gg::Algo: a plane is an array of LPB words.
//
// Two users:
//   * tests/hostsim/hostsim.cpp (CPU test simulator of the device algorithm);
//   * the thread-per-board rollout kernel variant (k_rollout_tpb in gg_kernels.cuh) for small boards, where
//     the LPB words live in registers of ONE thread: no shuffles, no ballots, 32 boards per warp.
// All loops have constant trip counts and static indices so that nvcc keeps the arrays in registers.
#pragma once
#include "gg_algo.cuh"

#ifndef GG_STAT_HFILL
#define GG_STAT_HFILL()
#endif
#ifndef GG_STAT_LOWEST
#define GG_STAT_LOWEST()
#endif

#if defined(__CUDACC__)
#define GG_UNROLL _Pragma("unroll")
#else
#define GG_UNROLL
#endif

namespace gg {

template <class G>
struct ArrayPlane {
    typename G::W w[G::LPB];
};
template <class G>
GG_HD ArrayPlane<G> operator|(ArrayPlane<G> a, const ArrayPlane<G>& b) {
    GG_UNROLL for (int j = 0; j < G::LPB; ++j) a.w[j] |= b.w[j];
    return a;
}
template <class G>
GG_HD ArrayPlane<G> operator&(ArrayPlane<G> a, const ArrayPlane<G>& b) {
    GG_UNROLL for (int j = 0; j < G::LPB; ++j) a.w[j] &= b.w[j];
    return a;
}

template <class G>
struct ArrayOps {
    typedef ArrayPlane<G> P;
    typedef typename G::W W;
    bool real;   // false: the thread has no board (tile tail) -> an all-zero board that never has a valid point
    GG_HD ArrayOps() : real(true) {}
    GG_HD explicit ArrayOps(bool r) : real(r) {}

    GG_HD P zero() const {
        P p;
        GG_UNROLL for (int j = 0; j < G::LPB; ++j) p.w[j] = 0;
        return p;
    }
    GG_HD P full() const {
        P p;
        GG_UNROLL for (int j = 0; j < G::LPB; ++j) p.w[j] = real ? G::rows_mask(G::rows_in_lane(j)) : W(0);
        return p;
    }
    GG_HD P andnot(P a, const P& b) const {
        GG_UNROLL for (int j = 0; j < G::LPB; ++j) a.w[j] &= ~b.w[j];
        return a;
    }
    GG_HD P east(P x) const {
        GG_UNROLL for (int j = 0; j < G::LPB; ++j) x.w[j] <<= 1;
        return x;
    }
    GG_HD P west(P x) const {
        GG_UNROLL for (int j = 0; j < G::LPB; ++j) x.w[j] >>= 1;
        return x;
    }
    GG_HD P south(const P& x) const {   // result[r] = x[r-1]
        P y;
        GG_UNROLL for (int j = 0; j < G::LPB; ++j) {
            const W in = G::RPL > 1 ? W(x.w[j] << (G::S % G::WB)) : W(0);
            const W prev = j ? W(x.w[j ? j - 1 : 0] >> ((G::RPL - 1) * G::S)) : W(0);
            y.w[j] = in | prev;
        }
        return y;
    }
    GG_HD P north(const P& x) const {   // result[r] = x[r+1]
        P y;
        GG_UNROLL for (int j = 0; j < G::LPB; ++j) {
            const W in = G::RPL > 1 ? W(x.w[j] >> (G::S % G::WB)) : W(0);
            // (rows 1.. of the next word land above the RPL row slots: outside every mask, like south's top row)
            const W next = j + 1 < G::LPB ? W(x.w[j + 1 < G::LPB ? j + 1 : j] << ((G::RPL - 1) * G::S)) : W(0);
            y.w[j] = in | next;
        }
        return y;
    }
    GG_HD P rev(P x) const {
        GG_UNROLL for (int j = 0; j < G::LPB; ++j) x.w[j] = w_rev(x.w[j]);
        return x;
    }
    GG_HD P hfill(P s, const P& m, const P& mrev) const {
        GG_STAT_HFILL();
        GG_UNROLL for (int j = 0; j < G::LPB; ++j) s.w[j] = w_hfill(s.w[j], m.w[j], mrev.w[j]);
        return s;
    }
    GG_HD bool any_board(const P& x) const {
        W a = 0;
        GG_UNROLL for (int j = 0; j < G::LPB; ++j) a |= x.w[j];
        return a != 0;
    }
    GG_HD bool any(const P& x) const { return any_board(x); }     // one board per thread: loops are thread-local
    GG_HD int popc(const P& x) const {
        int c = 0;
        GG_UNROLL for (int j = 0; j < G::LPB; ++j) c += w_popc(x.w[j]);
        return c;
    }
    GG_HD int count2(const P& x) const {
        const int c = popc(x);
        return c > 2 ? 2 : c;
    }
    GG_HD P lowest(const P& x) const {
        GG_STAT_LOWEST();
        P y;
        bool found = false;
        GG_UNROLL for (int j = 0; j < G::LPB; ++j) {
            y.w[j] = found ? W(0) : W(x.w[j] & (~x.w[j] + 1));
            found = found || x.w[j] != 0;
        }
        return y;
    }
    GG_HD P single(int pt) const {
        const int r = pt / G::N, c = pt - r * G::N;
        const int lj = r / G::RPL;
        const W bit = W(1) << ((r - lj * G::RPL) * G::S + c);
        P y;
        GG_UNROLL for (int j = 0; j < G::LPB; ++j) y.w[j] = j == lj ? bit : W(0);
        return y;
    }
    GG_HD P pick(bool c, const P& a, const P& b) const {
        P y;
        GG_UNROLL for (int j = 0; j < G::LPB; ++j) y.w[j] = c ? a.w[j] : b.w[j];
        return y;
    }
    GG_HD int kth_point(const P& x, int k) const {
        int pt = -1;
        GG_UNROLL for (int j = 0; j < G::LPB; ++j) {
            const int c = w_popc(x.w[j]);
            if (pt < 0 && k >= 0 && k < c) {
                const int bit = w_select(x.w[j], k);
                const int row = bit / G::S;
                pt = (j * G::RPL + row) * G::N + (bit - row * G::S);
            }
            k -= c;
        }
        return pt;
    }
};

}  // namespace gg
