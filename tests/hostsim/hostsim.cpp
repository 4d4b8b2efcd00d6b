// hostsim.cpp - CPU simulator of the DEVICE ALGORITHM (gymgo_b200/csrc/gg_algo.cuh).
//
// TEST INFRASTRUCTURE ONLY.  It instantiates the very same Algo<> template the sm_100a kernels use,
// with a plane held as an array of LPB words instead of one word per lane, so the CPU test-suite
// (this container has no GPU) can check the bit-level algorithm, the record layout and the sampler
// against the oracle on millions of positions before any GPU time is spent.  The product library
// never links or calls this file; gymgo_b200 has no CPU execution path.
#include <stdint.h>
#include <string.h>

static int g_stat_tag = 0;
static long g_stat_iters[8] = {0, 0, 0, 0, 0, 0, 0, 0};
// developer statistics (single-threaded test code): loop trips of the algorithm on the current board
static long g_stat_hfill = 0, g_stat_lowest = 0;
#define GG_STAT_TAG(k) g_stat_tag = (k);
#define GG_STAT_HFILL() (++g_stat_hfill, ++g_stat_iters[g_stat_tag & 7])
#define GG_STAT_LOWEST() (++g_stat_lowest)
#include "../../gymgo_b200/csrc/gg_array_ops.cuh"

namespace {

template <class G>
using HostPlane = gg::ArrayPlane<G>;
template <class G>
using HostOps = gg::ArrayOps<G>;

// ---- record <-> planes, dense <-> planes (the record layout of gg_algo.cuh Geo<>) ----
template <class G>
void rec_load(const uint32_t* rec, HostPlane<G>& b, HostPlane<G>& w, HostPlane<G>& i, uint32_t& flags) {
    HostPlane<G>* pl[3] = {&b, &w, &i};
    for (int k = 0; k < 3; ++k)
        for (int j = 0; j < G::LPB; ++j) {
            const uint32_t* p = rec + (k * G::LPB + j) * G::WW;
            pl[k]->w[j] = G::WW == 1 ? typename G::W(p[0]) : typename G::W(p[0] | (uint64_t(p[G::WW - 1]) << 32));
        }
    flags = rec[G::FLAGS_IDX];
}
template <class G>
void rec_store(uint32_t* rec, const HostPlane<G>& b, const HostPlane<G>& w, const HostPlane<G>& i, uint32_t flags) {
    const HostPlane<G>* pl[3] = {&b, &w, &i};
    memset(rec, 0, G::REC_BYTES);
    for (int k = 0; k < 3; ++k)
        for (int j = 0; j < G::LPB; ++j) {
            uint32_t* p = rec + (k * G::LPB + j) * G::WW;
            p[0] = uint32_t(pl[k]->w[j]);
            if (G::WW == 2) p[1] = uint32_t(uint64_t(pl[k]->w[j]) >> 32);
        }
    rec[G::FLAGS_IDX] = flags;
}
template <class G>
void dense_to_rec(const uint8_t* st, uint32_t* rec) {
    HostOps<G> o;
    HostPlane<G> pl[3] = {o.zero(), o.zero(), o.zero()};
    const int src[3] = {0, 1, 3};
    for (int k = 0; k < 3; ++k)
        for (int p = 0; p < G::NP; ++p)
            if (st[src[k] * G::NP + p]) pl[k] = pl[k] | o.single(p);
    uint32_t flags = 0;
    bool turn = false, pass = false, done = true;
    for (int p = 0; p < G::NP; ++p) {           // reference reads: max / max / all-ones (gogame.py:241-246,200-201,208-214)
        turn |= st[2 * G::NP + p] != 0;
        pass |= st[4 * G::NP + p] != 0;
        done &= st[5 * G::NP + p] != 0;
    }
    flags = (turn ? gg::FLAG_TURN : 0) | (pass ? gg::FLAG_PASS : 0) | (done ? gg::FLAG_DONE : 0);
    rec_store<G>(rec, pl[0], pl[1], pl[2], flags);
}
template <class G>
void rec_to_dense(const uint32_t* rec, uint8_t* st) {
    HostOps<G> o;
    HostPlane<G> pl[3];
    uint32_t flags;
    rec_load<G>(rec, pl[0], pl[1], pl[2], flags);
    const int dst[3] = {0, 1, 3};
    for (int k = 0; k < 3; ++k)
        for (int p = 0; p < G::NP; ++p) st[dst[k] * G::NP + p] = o.any_board(pl[k] & o.single(p));
    memset(st + 2 * G::NP, (flags & gg::FLAG_TURN) ? 1 : 0, G::NP);
    memset(st + 4 * G::NP, (flags & gg::FLAG_PASS) ? 1 : 0, G::NP);
    memset(st + 5 * G::NP, (flags & gg::FLAG_DONE) ? 1 : 0, G::NP);
}

template <class G>
struct Sim {
    static void layout(int* out) {
        out[0] = G::REC_BYTES; out[1] = G::LPB; out[2] = G::RPL; out[3] = G::WB; out[4] = G::BPW;
    }
    static void pack(const uint8_t* dense, int batch, uint32_t* recs) {
        for (int b = 0; b < batch; ++b) dense_to_rec<G>(dense + size_t(b) * 6 * G::NP, recs + size_t(b) * G::REC_W32);
    }
    static void unpack(const uint32_t* recs, int batch, uint8_t* dense) {
        for (int b = 0; b < batch; ++b) rec_to_dense<G>(recs + size_t(b) * G::REC_W32, dense + size_t(b) * 6 * G::NP);
    }
    static void step(const uint32_t* in, const int32_t* actions, int batch, uint32_t opts, uint32_t* out, uint8_t* status) {
        HostOps<G> o;
        for (int b = 0; b < batch; ++b) {
            HostPlane<G> bl, wh, iv;
            uint32_t flags;
            rec_load<G>(in + size_t(b) * G::REC_W32, bl, wh, iv, flags);
            int rc = gg::Algo<HostOps<G>>::step(o, G(), bl, wh, iv, flags, actions[b], opts);
            rec_store<G>(out + size_t(b) * G::REC_W32, bl, wh, iv, flags);
            if (status) status[b] = uint8_t(rc);
        }
    }
    static void areas(const uint32_t* in, int batch, int32_t* out) {
        HostOps<G> o;
        for (int b = 0; b < batch; ++b) {
            HostPlane<G> bl, wh, iv;
            uint32_t flags;
            rec_load<G>(in + size_t(b) * G::REC_W32, bl, wh, iv, flags);
            int ba, wa;
            gg::Algo<HostOps<G>>::areas(o, bl, wh, ba, wa);
            out[2 * b] = ba; out[2 * b + 1] = wa;
        }
    }
    // one fused rollout step as the device kernel does it: reset finished boards, sample, step
    static void rollout_step(uint32_t* recs, int batch, uint64_t seed, uint64_t board0, uint64_t t, int32_t* actions,
                             int32_t* stats = nullptr) {
        HostOps<G> o;
        for (int b = 0; b < batch; ++b) {
            g_stat_hfill = g_stat_lowest = 0;
            for (int k = 0; k < 8; ++k) g_stat_iters[k] = 0;
            HostPlane<G> bl, wh, iv;
            uint32_t flags;
            uint32_t* rec = recs + size_t(b) * G::REC_W32;
            rec_load<G>(rec, bl, wh, iv, flags);
            if (flags & gg::FLAG_DONE) { bl = wh = iv = o.zero(); flags = 0; }
            uint64_t gb = board0 + uint64_t(b);
            uint32_t rnd = gg::philox4x32_10(uint32_t(gb), uint32_t(gb >> 32), uint32_t(t), uint32_t(t >> 32),
                                             uint32_t(seed), uint32_t(seed >> 32));
            int a = gg::Algo<HostOps<G>>::sample_action(o, G(), iv, rnd);
            gg::Algo<HostOps<G>>::step(o, G(), bl, wh, iv, flags, a, 0u);
            rec_store<G>(rec, bl, wh, iv, flags);
            if (stats) {
                stats[8 * b] = int32_t(g_stat_hfill); stats[8 * b + 1] = int32_t(g_stat_lowest);
                for (int k = 0; k < 5; ++k) stats[8 * b + 2 + k] = int32_t(g_stat_iters[k]);
            }
            if (actions) actions[b] = a;
        }
    }
};

}  // namespace

#define GG_FOR_SIZES(X) X(2) X(3) X(4) X(5) X(6) X(7) X(8) X(9) X(10) X(11) X(12) X(13) X(14) X(15) X(16) X(17) X(18) X(19)

extern "C" {

int hs_layout(int n, int* out) {
    switch (n) {
#define X(NN) case NN: Sim<gg::Geo<NN>>::layout(out); return 0;
        GG_FOR_SIZES(X)
#undef X
    }
    return -1;
}
int hs_pack(int n, const uint8_t* dense, int batch, uint32_t* recs) {
    switch (n) {
#define X(NN) case NN: Sim<gg::Geo<NN>>::pack(dense, batch, recs); return 0;
        GG_FOR_SIZES(X)
#undef X
    }
    return -1;
}
int hs_unpack(int n, const uint32_t* recs, int batch, uint8_t* dense) {
    switch (n) {
#define X(NN) case NN: Sim<gg::Geo<NN>>::unpack(recs, batch, dense); return 0;
        GG_FOR_SIZES(X)
#undef X
    }
    return -1;
}
int hs_step(int n, const uint32_t* in, const int32_t* actions, int batch, uint32_t opts, uint32_t* out, uint8_t* status) {
    switch (n) {
#define X(NN) case NN: Sim<gg::Geo<NN>>::step(in, actions, batch, opts, out, status); return 0;
        GG_FOR_SIZES(X)
#undef X
    }
    return -1;
}
int hs_areas(int n, const uint32_t* in, int batch, int32_t* out) {
    switch (n) {
#define X(NN) case NN: Sim<gg::Geo<NN>>::areas(in, batch, out); return 0;
        GG_FOR_SIZES(X)
#undef X
    }
    return -1;
}
int hs_rollout_step(int n, uint32_t* recs, int batch, uint64_t seed, uint64_t board0, uint64_t t, int32_t* actions) {
    switch (n) {
#define X(NN) case NN: Sim<gg::Geo<NN>>::rollout_step(recs, batch, seed, board0, t, actions); return 0;
        GG_FOR_SIZES(X)
#undef X
    }
    return -1;
}
// developer statistics: stats[b][8] = {flood iterations, pocket-loop trips, iterations of flood kinds 0..4, -}
int hs_rollout_step_stats(int n, uint32_t* recs, int batch, uint64_t seed, uint64_t board0, uint64_t t, int32_t* actions,
                          int32_t* stats) {
    switch (n) {
#define X(NN) case NN: Sim<gg::Geo<NN>>::rollout_step(recs, batch, seed, board0, t, actions, stats); return 0;
        GG_FOR_SIZES(X)
#undef X
    }
    return -1;
}
uint32_t hs_philox(uint64_t board, uint64_t t, uint64_t seed) {
    return gg::philox4x32_10(uint32_t(board), uint32_t(board >> 32), uint32_t(t), uint32_t(t >> 32),
                             uint32_t(seed), uint32_t(seed >> 32));
}

}  // extern "C"
