// gg_algo.cuh - the Go rules as a branch-light bitboard algebra, shared verbatim by
//   * the sm_100a kernels (gg_kernels.cu): a board is spread over LPB adjacent lanes of a warp,
//     one machine word per lane per plane, cross-lane traffic by __shfl_sync/__ballot_sync;
//   * the host simulator (tests/hostsim/hostsim.cpp): the same code with the LPB words in an array,
//     used ONLY by CPU tests to check this algorithm against the oracle before GPU time is spent.
//
// What it restates (reference file:line, behaviour only - see SURVEY.md Appendix A):
//   step()            gym_go/gogame.py:34-87  (+ state_utils.py:159-180 capture, :214-223 adjacency)
//   invalid_mask()    gym_go/state_utils.py:24-83
//   areas()           gym_go/gogame.py:275-300
//
// Layout of one plane of one board ("guard-column bitboard"):
//   row r, column c  ->  lane j = r / RPL, bit (r % RPL) * S + c   with row stride S = N + 1.
//   Bit N of every row slot is a guard bit that is 0 in every plane, so
//     - east/west neighbours are plain 1-bit shifts (no column masks on the data path),
//     - a whole-row flood is ONE addition: carries run through consecutive mask bits and die in
//       the guard bit (hfill below); the opposite direction uses the bit-reversed word.
//   north/south neighbours are a shift by S inside the word plus one shuffle with the adjacent lane.
//
// Every collective (any/shuffle/ballot) is executed by all 32 lanes of the warp, so the code below is
// predicated per board instead of branching per board: a board that passes, or whose move is refused,
// runs the move path with an empty "move" plane, which makes every flood a no-op for it.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define GG_HD __host__ __device__ __forceinline__
#else
#define GG_HD inline
#endif

// developer hook: the host simulator can tag which flood is running (tools/algo_stats.py); a no-op on the device
#ifndef GG_STAT_TAG
#define GG_STAT_TAG(k)
#endif

namespace gg {

enum : uint32_t { FLAG_TURN = 1u, FLAG_PASS = 2u, FLAG_DONE = 4u };
enum : int { ST_OK = 0, ST_INVALID_MOVE = 1, ST_OUT_OF_RANGE = 2, ST_GAME_OVER = 3 };
// step option bits (also the `flags` argument of gg_step in include/gymgo_b200.h)
enum : uint32_t { OPT_CANONICAL = 1u, OPT_REFUSE_DONE = 2u, OPT_AUTO_RESET = 4u, OPT_RESET_SKIPS_ACTION = 8u,
                  OPT_KERNEL_LANES = 16u, OPT_KERNEL_THREAD = 32u };   // (kernel selectors: not rule options)

constexpr int cdiv(int a, int b) { return (a + b - 1) / b; }
constexpr int default_wordbits(int n) { return n <= 9 ? 32 : 64; }

template <int WB> struct WordOf;
template <> struct WordOf<32> { typedef uint32_t type; };
template <> struct WordOf<64> { typedef uint64_t type; };

// Geometry of the packed record for board size N_.
template <int N_, int WORDBITS = default_wordbits(N_)>
struct Geo {
    static constexpr int N = N_, S = N_ + 1, NP = N_ * N_, A = N_ * N_ + 1;
    static constexpr int WB = WORDBITS;
    static constexpr int RPL_MAX = WB / S;                 // rows that fit one word
    static constexpr int LPB = cdiv(N, RPL_MAX);           // lanes (words) per board per plane
    static constexpr int RPL = cdiv(N, LPB);               // rows per lane, balanced
    static constexpr int BPW = 32 / LPB;                   // boards per warp
    static constexpr int WW = WB / 32;                     // 32-bit words per lane word
    static constexpr int PLANE_W32 = LPB * WW;
    static constexpr int FLAGS_IDX = 3 * PLANE_W32;        // record = black | white | invalid | flags
    static constexpr int REC_W32 = (FLAGS_IDX + 1 + 3) / 4 * 4;   // padded to 16 B for bulk copies
    static constexpr int REC_BYTES = REC_W32 * 4;
    typedef typename WordOf<WB>::type W;
    static_assert(N_ >= 2 && RPL_MAX >= 1 && RPL * S <= WB && LPB <= 32, "unsupported board size");

    static GG_HD constexpr W row_bits() { return (W(1) << N) - 1; }
    static GG_HD W rows_mask(int rows) {
        W m = 0;
        for (int i = 0; i < RPL; ++i)
            if (i < rows) m |= row_bits() << (i * S);
        return m;
    }
    static GG_HD int rows_in_lane(int j) {
        int r = N - j * RPL;
        return r < 0 ? 0 : (r > RPL ? RPL : r);
    }
};

// ---------------------------------------------------------------- word helpers (host + device)
GG_HD uint32_t w_rev(uint32_t x) {
#if defined(__CUDA_ARCH__)
    return __brev(x);
#else
    x = (x >> 16) | (x << 16);
    x = ((x & 0xff00ff00u) >> 8) | ((x & 0x00ff00ffu) << 8);
    x = ((x & 0xf0f0f0f0u) >> 4) | ((x & 0x0f0f0f0fu) << 4);
    x = ((x & 0xccccccccu) >> 2) | ((x & 0x33333333u) << 2);
    x = ((x & 0xaaaaaaaau) >> 1) | ((x & 0x55555555u) << 1);
    return x;
#endif
}
GG_HD uint64_t w_rev(uint64_t x) {
#if defined(__CUDA_ARCH__)
    return __brevll(x);
#else
    return (uint64_t(w_rev(uint32_t(x))) << 32) | w_rev(uint32_t(x >> 32));
#endif
}
GG_HD int w_popc(uint32_t x) {
#if defined(__CUDA_ARCH__)
    return __popc(x);
#else
    return __builtin_popcount(x);
#endif
}
GG_HD int w_popc(uint64_t x) {
#if defined(__CUDA_ARCH__)
    return __popcll(x);
#else
    return __builtin_popcountll(x);
#endif
}
// index of the lowest set bit (x != 0)
GG_HD int w_ctz(uint32_t x) {
#if defined(__CUDA_ARCH__)
    return __ffs(int(x)) - 1;
#else
    return __builtin_ctz(x);
#endif
}
GG_HD int w_ctz(uint64_t x) {
#if defined(__CUDA_ARCH__)
    return __ffsll((long long)x) - 1;
#else
    return __builtin_ctzll(x);
#endif
}
// position of the k-th (0-based) set bit of x; requires k < popc(x).  Branch-light binary search on
// popcounts of the low halves (5 / 6 rounds) instead of clearing k bits one by one.
GG_HD int w_select(uint32_t x, int k) {
    int pos = 0;
#define GG_SEL_STEP(BITS)                                   \
    {                                                       \
        const int c = w_popc(uint32_t(x & ((1u << BITS) - 1u))); \
        const bool up = k >= c;                             \
        k -= up ? c : 0;                                    \
        pos += up ? BITS : 0;                               \
        x = up ? (x >> BITS) : x;                           \
    }
    GG_SEL_STEP(16) GG_SEL_STEP(8) GG_SEL_STEP(4) GG_SEL_STEP(2) GG_SEL_STEP(1)
#undef GG_SEL_STEP
    return pos;
}
GG_HD int w_select(uint64_t x, int k) {
    const int c = w_popc(uint32_t(x));
    return k >= c ? 32 + w_select(uint32_t(x >> 32), k - c) : w_select(uint32_t(x), k);
}

// Row flood inside one word: every maximal run of consecutive bits of `m` that contains a bit of
// `s` (s subset of m) becomes fully set.  a = m + s ripples a carry from each seed to the top of its run
// (it dies in the 0 bit above: guard bit / word end), so inside m the bits of a run at and above its lowest seed
// come out as ~a (except further seeds); the bit-reversed word does the other direction: b = rev(mrev + rev(s)).
// A run bit is reached iff it is flipped in one of the two sums: m & ~(a & b), plus the seeds themselves (a seed
// with seeds on both sides is flipped back in both).  Two additions, two reversals, two logic operations.
template <class W>
GG_HD W w_hfill(W s, W m, W mrev) {
    const W a = m + s;
    const W b = w_rev(W(mrev + w_rev(s)));
    return (m & ~(a & b)) | s;
}

// Philox4x32-10 (Salmon et al., SC'11) - counter-based, so a board's random stream depends only on
// (seed, global board index, step index), never on how boards are sharded over GPUs.  Returns word 0.
GG_HD uint32_t philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
    for (int r = 0; r < 10; ++r) {
        const uint64_t p0 = uint64_t(0xD2511F53u) * c0;
        const uint64_t p1 = uint64_t(0xCD9E8D57u) * c2;
        const uint32_t n0 = uint32_t(p1 >> 32) ^ c1 ^ k0;
        const uint32_t n2 = uint32_t(p0 >> 32) ^ c3 ^ k1;
        c1 = uint32_t(p1); c3 = uint32_t(p0); c0 = n0; c2 = n2;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    return c0;
}

// =================================================================================================
// The algorithm, written against an `Ops` policy:
//   typedef P                      plane (device: one word; host: LPB words)
//   P zero(), full()               empty plane / all real points of this board
//   P east(P) west(P)              1-column moves   (result must still be masked with full())
//   P south(P) north(P)            1-row moves      (south: result[r] = x[r-1]; north: x[r+1]; may leave bits above
//                                  the word's row slots: like east/west, mask the result with full() or a plane)
//   P rev(P), hfill(s, m, mrev)    per-word bit reversal / row flood
//   bool any(P)                    LOOP CONTROL ONLY: true if any board sharing my warp has a bit
//   bool any_board(P)              this board has a bit
//   int  count2(P)                 min(popcount over this board, 2)
//   int  popc(P)                   popcount over this board
//   P lowest(P)                    only the lowest set point of this board (empty if none)
//   P single(int pt)               plane with point pt (row-major index) set; pt must be in range
//   P pick(bool c, P a, P b)       c ? a : b
//   int kth_point(P x, int k)      row-major index of the k-th (0-based) set point of this board, k < popc(x)
// =================================================================================================
template <class O>
struct Algo {
    typedef typename O::P P;

    static GG_HD P nbrs(const O& o, P x) {          // 4-neighbourhood, in-board only, centre excluded
        return (o.east(x) | o.west(x) | o.south(x) | o.north(x)) & o.full();
    }

    // stones of `mask` 4-connected (inside mask) to `seed`; seed must be a subset of mask; mrev = rev(mask)
    static GG_HD P flood(const O& o, P seed, P mask, P mrev) {
        P x = seed;
        for (;;) {
            x = o.hfill(x, mask, mrev);
            P v = o.andnot((o.south(x) | o.north(x)) & mask, x);
            if (!o.any(v)) break;
            x = x | v;
        }
        return x;
    }
    static GG_HD P flood(const O& o, P seed, P mask) { return flood(o, seed, mask, o.rev(mask)); }

    // two independent floods advanced together: one vote per iteration instead of two and two independent
    // dependency chains for the scheduler to interleave (the iteration count is the larger of the two)
    static GG_HD void flood2(const O& o, P& x1, P m1, P& x2, P m2) {
        const P r1 = o.rev(m1), r2 = o.rev(m2);
        for (;;) {
            x1 = o.hfill(x1, m1, r1);
            x2 = o.hfill(x2, m2, r2);
            const P v1 = o.andnot((o.south(x1) | o.north(x1)) & m1, x1);
            const P v2 = o.andnot((o.south(x2) | o.north(x2)) & m2, x2);
            if (!o.any(v1 | v2)) break;
            x1 = x1 | v1;
            x2 = x2 | v2;
        }
    }

    // INVD plane for the player whose stones are `nxt` (to move), `oth` = the player who just moved.
    // invalid(p) = occupied | ko | ( no empty neighbour
    //                                & not adjacent to an `oth` group with exactly one liberty
    //                                & not adjacent to a  `nxt` group with two or more liberties )
    // Only "pockets" (empty points without an empty neighbour) can fall in the third class, and only the
    // groups next to them matter.  Most of those are settled without visiting them one by one:
    //   * a group owning a liberty that is not a pocket, or a stone that touches two empty points, has
    //     >= 2 liberties: ONE flood per colour ("big") finds all of them;
    //   * of what is left every stone touches at most one empty point, so a stone without a same-coloured
    //     neighbour is a one-stone group in atari: bit-parallel, no flood;
    //   * the remaining groups (all their liberties are pockets) are flooded one per trip.
    static GG_HD P invalid_mask(const O& o, P nxt, P oth, P ko) {
        const P occ = nxt | oth;
        const P empty = o.andnot(o.full(), occ);
        const P ee = o.east(empty), ew = o.west(empty), es = o.south(empty), en = o.north(empty);
        const P pockets = o.andnot(empty, ee | ew | es | en);
        P bad = pockets;                               // pockets not yet shown to be playable
        if (o.any(pockets)) {
            const P open = o.andnot(empty, pockets);   // liberties that have an empty neighbour
            const P two = (ee & ew) | (es & en) | ((ee | ew) & (es | en));   // touches >= 2 empty points
            const P healthy = nbrs(o, open) | two;
            GG_STAT_TAG(2)
            P big_nxt = nxt & healthy, big_oth = oth & healthy;             // -> groups with >= 2 liberties (sufficient)
            flood2(o, big_nxt, nxt, big_oth, oth);
            bad = o.andnot(bad, nbrs(o, big_nxt));     // own group with >= 2 liberties: safe
            const P rest_nxt = o.andnot(nxt, big_nxt);
            const P rest_oth = o.andnot(oth, big_oth);
            if (o.any(bad)) {
                const P adj = nbrs(o, bad);
                const P cand_nxt = rest_nxt & adj, cand_oth = rest_oth & adj;
                const P lone_nxt = o.andnot(cand_nxt, nbrs(o, nxt));       // one stone, one liberty: no help
                const P lone_oth = o.andnot(cand_oth, nbrs(o, oth));       // one stone in atari: capturable
                bad = o.andnot(bad, nbrs(o, lone_oth));
                P todo = (o.andnot(cand_nxt, lone_nxt) | o.andnot(cand_oth, lone_oth)) & nbrs(o, bad);
                const P rrev_nxt = o.rev(rest_nxt), rrev_oth = o.rev(rest_oth);
                while (o.any(todo)) {
                    const P s = o.lowest(todo);
                    const bool mine = o.any_board(s & nxt);
                    GG_STAT_TAG(4)
                    const P grp = flood(o, s, o.pick(mine, rest_nxt, rest_oth), o.pick(mine, rrev_nxt, rrev_oth));
                    const P libs = nbrs(o, grp) & empty;
                    const int nl = o.count2(libs);
                    const bool playable = mine ? (nl >= 2) : (nl == 1);   // stays alive / captures
                    bad = o.andnot(bad, o.pick(playable, libs, o.zero()));
                    todo = o.andnot(todo, grp) & nbrs(o, bad);
                }
            }
        }
        return occ | bad | ko;
    }

    // Stones of `opp` that belong to groups touching a point of `touch` and have no liberty left (update_pieces,
    // state_utils.py:159-180).  A dead group consists only of "enclosed" stones (no empty neighbour of their own), so
    // only the enclosed stones on `touch` are flooded, and only through enclosed stones: in open positions this is
    // empty or tiny.  Such a component is alive iff it touches an opponent stone that is not enclosed.
    static GG_HD P captured(const O& o, P opp, P own, P touch) {
        const P empty_now = o.andnot(o.full(), own | opp);
        const P enclosed = o.andnot(opp, nbrs(o, empty_now));
        const P seeds = touch & enclosed;
        P dead = o.zero();
        if (o.any(seeds)) {
            GG_STAT_TAG(0)
            const P grp = flood(o, seeds, enclosed);               // enclosed components touching the move
            GG_STAT_TAG(1)
            const P alive = flood(o, grp & nbrs(o, o.andnot(opp, enclosed)), grp);
            dead = o.andnot(grp, alive);
        }
        return dead;
    }

    // One ply.  Planes and flags are updated in place when the returned status is ST_OK and left
    // untouched otherwise.  `action` in [0, N*N] (N*N = pass).
    template <class G>
    static GG_HD int step(const O& o, G, P& black, P& white, P& invd, uint32_t& flags, int action, uint32_t opts) {
        const bool white_to_move = (flags & FLAG_TURN) != 0;
        const bool in_range = action >= 0 && action <= G::NP;
        const bool is_pass = action == G::NP;
        const bool refused_done = (opts & OPT_REFUSE_DONE) && (flags & FLAG_DONE);
        P m = o.zero();
        if (in_range && !is_pass) m = o.single(action);            // per-board data, no collective inside
        const bool illegal = o.any_board(m & invd);
        int status = ST_OK;
        if (!in_range) status = ST_OUT_OF_RANGE;
        else if (refused_done) status = ST_GAME_OVER;
        else if (illegal) status = ST_INVALID_MOVE;
        if (status != ST_OK) m = o.zero();

        P own = o.pick(white_to_move, white, black);
        P opp = o.pick(white_to_move, black, white);
        own = own | m;
        // neighbours of the move; "surrounded" is judged before captures (state_utils.py:220-221)
        const P nb = nbrs(o, m);
        // (collectives are never placed behind a short-circuit: every lane of the warp must reach them)
        const bool placed = o.any_board(m);
        const bool gap = o.any_board(o.andnot(nb, opp));
        const bool hemmed = placed && !gap;
        // captures (state_utils.py:159-180), then the ko point: one group of one stone died and the move is hemmed in
        const P dead = captured(o, opp, own, nb);
        P ko = o.zero();
        if (o.any(dead)) {                                                 // uniform over the boards that share a vote
            opp = o.andnot(opp, dead);
            const int ndead = o.count2(dead);
            ko = o.pick(hemmed && ndead == 1, dead, o.zero());
        }
        // mask for the player who moves next (= opp colour), also recomputed on a pass (ko expires)
        const P new_invd = invalid_mask(o, opp, own, ko);

        uint32_t nf = flags ^ FLAG_TURN;
        if (is_pass) {
            if (flags & FLAG_PASS) nf |= FLAG_DONE;
            nf |= FLAG_PASS;
        } else {
            nf &= ~FLAG_PASS;
        }
        P nb_black = o.pick(white_to_move, opp, own);
        P nb_white = o.pick(white_to_move, own, opp);
        if ((opts & OPT_CANONICAL) && (nf & FLAG_TURN)) {          // gogame.py:313-321
            P t = nb_black; nb_black = nb_white; nb_white = t;
            nf &= ~FLAG_TURN;
        }
        if (status == ST_OK) { black = nb_black; white = nb_white; invd = new_invd; flags = nf; }
        return status;
    }

    // Uniform choice among the valid actions incl. pass (reference go_env.py:78-81), driven by one
    // 32-bit random word: k = floor(rnd * count / 2^32) picks the k-th valid action in ascending
    // action order; the pass (index N*N) is the last one.
    template <class G>
    static GG_HD int sample_action(const O& o, G, P invd, uint32_t rnd) {
        const P ok = o.andnot(o.full(), invd);
        const int count = o.popc(ok) + 1;
        const int k = int((uint64_t(rnd) * uint64_t(count)) >> 32);
        const int pt = o.kth_point(ok, k < count - 1 ? k : 0);
        return k == count - 1 ? G::NP : pt;
    }

    // Tromp-Taylor areas: stones + empty points reachable (through empties) from one colour only.
    static GG_HD void areas(const O& o, P black, P white, int& black_area, int& white_area) {
        const P empty = o.andnot(o.full(), black | white);
        P rb = nbrs(o, black) & empty, rw = nbrs(o, white) & empty;
        flood2(o, rb, empty, rw, empty);
        black_area = o.popc(black | o.andnot(rb, rw));
        white_area = o.popc(white | o.andnot(rw, rb));
    }
};

}  // namespace gg
