// gg_host.cpp - HOST codec behind gg_host_unpack: packed records in host memory -> dense [B,6,N,N] in host memory.
//
// Why it exists: a consumer that lives in host memory and wants float32 observations is PCIe-bound when the dense
// tensor crosses the bus (127.7 MB per 9x9 x 65,536 step at ~55 GB/s).  The packed records are 40x smaller, so the
// cheaper route is records over PCIe + expansion next to the CPU - provided the expansion runs at memory speed.
// This file makes it do so: a persistent worker pool, the batch cut into 32-board chunks (whole 64-bit stream words,
// 64-byte aligned output for every dtype), per chunk
//   1. the chunk's dense bits in final element order in a small stack buffer (guard bits dropped with PEXT,
//      constant planes as runs of ones) - ~12 appends per 9x9 board;
//   2. expansion with AVX-512 mask moves: one k-register load + one masked move + one NON-TEMPORAL 64-byte store per
//      16 floats / 32 halves / 64 bytes / 8 doubles (the output is write-only and far larger than the caches).
// A scalar path serves CPUs without AVX-512BW/BMI2 and misaligned outputs.
// No Go rules run here: it is gg_unpack's layout (include/gymgo_b200.h "Packed record"), nothing else.
#include <immintrin.h>
#include <stdint.h>
#include <string.h>

#include <atomic>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

#include "gg_host.h"

namespace gg {
namespace {

// ------------------------------------------------------------------------------------------ worker pool
// Persistent threads (created on first use, grown on demand, never joined: the library has no tear-down entry point).
// One job at a time (callers serialise on run_mu_).  A job is finished when all its ITEMS are done - not when every
// helper has reported - so a helper that wakes up late costs nothing: the caller and the punctual helpers take its
// share.  Tickets carry the job's epoch, so a late helper can never take an item of a later job with a stale snapshot.
// Helpers block on a condition variable between jobs: polling for the next job was measured and bought nothing in a
// stepping loop while it slowed every other host thread down (profiles/r02_host_codec_probe.json).
class Pool {
  public:
    void run(int workers, int64_t items, const std::function<void(int64_t)>& fn) {
        std::lock_guard<std::mutex> only_one(run_mu_);
        if (workers <= 1 || items <= 1) {
            for (int64_t i = 0; i < items; ++i) fn(i);
            return;
        }
        Job job;
        {
            std::lock_guard<std::mutex> lk(mu_);
            while (int(threads_.size()) < workers - 1) threads_.emplace_back([this, id = int(threads_.size())] { loop(id); });
            job.fn = &fn;
            job.items = items;
            job.helpers = workers - 1;
            job.epoch = job_.epoch + 1;
            job_ = job;
            done_.store(0, std::memory_order_relaxed);
            next_.store((job.epoch & EPOCH_MASK) << INDEX_BITS, std::memory_order_relaxed);
            epoch_.store(job.epoch, std::memory_order_release);
        }
        cv_.notify_all();
        work(job);
        for (int spin = 0; done_.load(std::memory_order_acquire) != items; ++spin) {   // items in flight elsewhere
            if (spin < 4096) _mm_pause();
            else std::this_thread::yield();
        }
    }

  private:
    struct Job {
        const std::function<void(int64_t)>* fn = nullptr;
        int64_t items = 0;
        int helpers = 0;
        uint64_t epoch = 0;
    };
    static constexpr int INDEX_BITS = 40;
    static constexpr uint64_t INDEX_MASK = (uint64_t(1) << INDEX_BITS) - 1, EPOCH_MASK = (uint64_t(1) << 24) - 1;

    void work(const Job& j) {
        uint64_t v = next_.load(std::memory_order_relaxed);
        for (;;) {
            if ((v >> INDEX_BITS) != (j.epoch & EPOCH_MASK) || int64_t(v & INDEX_MASK) >= j.items) return;
            if (!next_.compare_exchange_weak(v, v + 1, std::memory_order_relaxed)) continue;      // v reloaded
            (*j.fn)(int64_t(v & INDEX_MASK));                      // the caller waits for this item: fn is alive
            done_.fetch_add(1, std::memory_order_release);
            v = next_.load(std::memory_order_relaxed);
        }
    }
    void loop(int id) {
        uint64_t seen = 0;
        for (;;) {
            Job j;
            {
                std::unique_lock<std::mutex> lk(mu_);
                cv_.wait(lk, [&] { return epoch_.load(std::memory_order_acquire) != seen; });
                j = job_;                                          // published under this lock
            }
            seen = j.epoch;
            if (id < j.helpers) work(j);                           // else: this job wants fewer workers than the pool holds
        }
    }
    std::mutex run_mu_, mu_;
    std::condition_variable cv_;
    std::vector<std::thread> threads_;
    Job job_;
    std::atomic<uint64_t> next_{0}, epoch_{0};
    std::atomic<int64_t> done_{0};
};
Pool& pool() {
    static Pool* p = new Pool;                              // leaked on purpose (threads outlive static destructors)
    return *p;
}

// ------------------------------------------------------------------------------------------ bit stream of a chunk
constexpr int CHUNK = 32;                                   // boards per chunk: 32 * 6 * N*N bits = 3*N*N whole words
constexpr int MAX_WORDS = 3 * 19 * 19 + 3;

struct Geo {
    int n, np, s, lpb, rpl, word_bytes, plane_bytes, rec_bytes;
    uint64_t rows_mask[8];                                  // mask of the real points of a word holding 0..7 rows (rpl <= 5)
};

// Bit appender that keeps the word under construction in a register (no read-modify-write chain through memory) and
// is branch-free: every append stores the current word (it is simply stored again until it is full) and steps to
// the next one when it fills up; finish() stores the last, partial word.  The spill into the next word is (v >> 1) >> (63 - fill): 0 when nothing spills.
struct BitSink {
    uint64_t* out;
    uint64_t acc = 0;
    int fill = 0;
    explicit BitSink(uint64_t* o) : out(o) {}
    inline void put(uint64_t v, int nbits) {                       // 1 <= nbits <= 64, no bits of v at or above nbits
        acc |= v << fill;
        *out = acc;
        const int nf = fill + nbits;
        const bool wrap = nf >= 64;
        const uint64_t spill = (v >> 1) >> (63 - fill);
        acc = wrap ? spill : acc;
        out += wrap;
        fill = nf & 63;
    }
    inline void finish() { *out = acc; }                           // the word under construction (stream has a spare word)
    inline void put_run(uint64_t on, int nbits) {                  // nbits copies of one bit (on = 0 or ~0)
        for (; nbits >= 64; nbits -= 64) put(on, 64);
        if (nbits) put(on & ((uint64_t(1) << nbits) - 1), nbits);
    }
};
inline uint64_t compact_scalar(uint64_t w, int rows, int n, int s) {
    uint64_t out = 0;
    const uint64_t row = (uint64_t(1) << n) - 1;
    for (int i = 0; i < rows; ++i) out |= ((w >> (i * s)) & row) << (i * n);
    return out;
}

// PEXT through inline assembly: usable from code compiled without -mbmi2 (only ever executed when the CPU has BMI2)
inline uint64_t pext64(uint64_t w, uint64_t mask) {
    uint64_t out;
    asm("pext %2, %1, %0" : "=r"(out) : "r"(w), "r"(mask));
    return out;
}

// WORD = uint32_t / uint64_t: the record's word type (fixed-size loads instead of a run-time memcpy length)
template <bool BMI2, class WORD>
inline void chunk_bits(const Geo& g, const uint8_t* rec, int boards, uint64_t* stream) {
    BitSink sink(stream);
    const int full_words = g.n / g.rpl, tail_rows = g.n - full_words * g.rpl;      // words holding rpl rows, rows of the last
    const uint64_t full_mask = g.rows_mask[g.rpl], tail_mask = g.rows_mask[tail_rows];
    const int full_bits = g.rpl * g.n, tail_bits = tail_rows * g.n;
    for (int b = 0; b < boards; ++b) {
        const uint8_t* r = rec + size_t(b) * g.rec_bytes;
        uint32_t flags;
        memcpy(&flags, r + 3 * g.plane_bytes, 4);
        for (int ch = 0; ch < 6; ++ch) {
            if (ch == 2 || ch == 4 || ch == 5) {            // constant planes: turn / previous pass / game over
                const int bit = ch == 2 ? 0 : (ch == 4 ? 1 : 2);
                sink.put_run(uint64_t(0) - uint64_t((flags >> bit) & 1u), g.np);
                continue;
            }
            const uint8_t* plane = r + (ch == 3 ? 2 : ch) * g.plane_bytes;
            for (int j = 0; j < full_words; ++j) {
                WORD w;
                memcpy(&w, plane + j * sizeof(WORD), sizeof(WORD));
                sink.put(BMI2 ? pext64(w, full_mask) : compact_scalar(w, g.rpl, g.n, g.s), full_bits);
            }
            if (tail_rows) {
                WORD w;
                memcpy(&w, plane + full_words * sizeof(WORD), sizeof(WORD));
                sink.put(BMI2 ? pext64(w, tail_mask) : compact_scalar(w, tail_rows, g.n, g.s), tail_bits);
            }
        }
    }
    sink.finish();
}

// ------------------------------------------------------------------------------------------ expansion
template <class T>
void expand_scalar(const uint64_t* stream, size_t first, size_t count, T one, T* out) {
    for (size_t i = first; i < first + count; ++i) out[i] = ((stream[i >> 6] >> (i & 63)) & 1) ? one : T(0);
}

#if defined(__GNUC__)
#define GG_AVX512 __attribute__((target("avx512f,avx512bw,avx512vl,avx512dq,bmi2")))
#else
#define GG_AVX512
#endif

// each handles whole 64-bit stream words [0, words); `out` is 64-byte aligned when NT
template <bool NT>
GG_AVX512 void expand_f32(const uint64_t* stream, size_t words, float* out) {
    const __m512 ones = _mm512_set1_ps(1.0f);
    for (size_t w = 0; w < words; ++w) {
        uint64_t bits = stream[w];
        float* o = out + w * 64;
        for (int q = 0; q < 4; ++q, bits >>= 16) {
            const __m512 v = _mm512_maskz_mov_ps(__mmask16(bits), ones);
            if (NT) _mm512_stream_ps(o + q * 16, v);
            else _mm512_storeu_ps(o + q * 16, v);
        }
    }
}
template <bool NT>
GG_AVX512 void expand_16(const uint64_t* stream, size_t words, uint16_t one, uint16_t* out) {
    const __m512i ones = _mm512_set1_epi16(short(one));
    for (size_t w = 0; w < words; ++w) {
        uint64_t bits = stream[w];
        uint16_t* o = out + w * 64;
        for (int q = 0; q < 2; ++q, bits >>= 32) {
            const __m512i v = _mm512_maskz_mov_epi16(__mmask32(bits), ones);
            if (NT) _mm512_stream_si512(reinterpret_cast<__m512i*>(o + q * 32), v);
            else _mm512_storeu_si512(o + q * 32, v);
        }
    }
}
template <bool NT>
GG_AVX512 void expand_u8(const uint64_t* stream, size_t words, uint8_t* out) {
    const __m512i ones = _mm512_set1_epi8(1);
    for (size_t w = 0; w < words; ++w) {
        const __m512i v = _mm512_maskz_mov_epi8(__mmask64(stream[w]), ones);
        if (NT) _mm512_stream_si512(reinterpret_cast<__m512i*>(out + w * 64), v);
        else _mm512_storeu_si512(out + w * 64, v);
    }
}
template <bool NT>
GG_AVX512 void expand_f64(const uint64_t* stream, size_t words, double* out) {
    const __m512d ones = _mm512_set1_pd(1.0);
    for (size_t w = 0; w < words; ++w) {
        uint64_t bits = stream[w];
        double* o = out + w * 64;
        for (int q = 0; q < 8; ++q, bits >>= 8) {
            const __m512d v = _mm512_maskz_mov_pd(__mmask8(bits), ones);
            if (NT) _mm512_stream_pd(o + q * 8, v);
            else _mm512_storeu_pd(o + q * 8, v);
        }
    }
}

bool cpu_has_avx512() {
#if defined(__GNUC__) && (defined(__x86_64__) || defined(__i386__))
    static const bool ok = __builtin_cpu_supports("avx512f") && __builtin_cpu_supports("avx512bw") &&
                           __builtin_cpu_supports("avx512vl") && __builtin_cpu_supports("avx512dq") &&
                           __builtin_cpu_supports("bmi2");
    return ok;
#else
    return false;
#endif
}

int elem_bytes(int dtype) { return dtype == HOST_U8 ? 1 : (dtype == HOST_F32 ? 4 : (dtype == HOST_F64 ? 8 : 2)); }

// one chunk of `boards` (<= CHUNK) boards starting at record pointer `rec`, dense output at `out`
void do_chunk(const Geo& g, const uint8_t* rec, int boards, int dtype, uint16_t one16, void* out, bool simd, bool nt) {
    uint64_t stream[MAX_WORDS];
    const size_t elems = size_t(boards) * 6 * g.np;
    if (g.word_bytes == 4) simd ? chunk_bits<true, uint32_t>(g, rec, boards, stream) : chunk_bits<false, uint32_t>(g, rec, boards, stream);
    else simd ? chunk_bits<true, uint64_t>(g, rec, boards, stream) : chunk_bits<false, uint64_t>(g, rec, boards, stream);
    const size_t words = simd ? elems / 64 : 0;             // whole words go through the vector path
    if (words) {
        switch (dtype) {
            case HOST_F32: nt ? expand_f32<true>(stream, words, static_cast<float*>(out)) : expand_f32<false>(stream, words, static_cast<float*>(out)); break;
            case HOST_U8: nt ? expand_u8<true>(stream, words, static_cast<uint8_t*>(out)) : expand_u8<false>(stream, words, static_cast<uint8_t*>(out)); break;
            case HOST_F64: nt ? expand_f64<true>(stream, words, static_cast<double*>(out)) : expand_f64<false>(stream, words, static_cast<double*>(out)); break;
            default: nt ? expand_16<true>(stream, words, one16, static_cast<uint16_t*>(out)) : expand_16<false>(stream, words, one16, static_cast<uint16_t*>(out)); break;
        }
    }
    const size_t first = words * 64, rest = elems - first;  // ragged last chunk / scalar path
    if (rest) {
        switch (dtype) {
            case HOST_F32: expand_scalar<float>(stream, first, rest, 1.0f, static_cast<float*>(out)); break;
            case HOST_U8: expand_scalar<uint8_t>(stream, first, rest, uint8_t(1), static_cast<uint8_t*>(out)); break;
            case HOST_F64: expand_scalar<double>(stream, first, rest, 1.0, static_cast<double*>(out)); break;
            default: expand_scalar<uint16_t>(stream, first, rest, one16, static_cast<uint16_t*>(out)); break;
        }
    }
}

}  // namespace

const char* host_unpack_path() { return cpu_has_avx512() ? "avx512 mask-move + pext, non-temporal stores" : "scalar"; }

void host_unpack(const uint8_t* rec, int64_t batch, int n, int lpb, int rpl, int wordbits, int rec_bytes, int dtype,
                 void* dense, int threads) {
    if (batch <= 0) return;
    Geo g;
    g.n = n;
    g.np = n * n;
    g.s = n + 1;
    g.lpb = lpb;
    g.rpl = rpl;
    g.word_bytes = wordbits / 8;
    g.plane_bytes = lpb * g.word_bytes;
    g.rec_bytes = rec_bytes;
    for (int rows = 0; rows < 8; ++rows) {
        uint64_t m = 0;
        for (int i = 0; i < rows && i < rpl; ++i) m |= ((uint64_t(1) << n) - 1) << (i * g.s);
        g.rows_mask[rows] = m;
    }
    const uint16_t one16 = dtype == HOST_BF16 ? uint16_t(0x3F80) : uint16_t(0x3C00);
    const bool simd = cpu_has_avx512();
    const bool nt = simd && (reinterpret_cast<uintptr_t>(dense) & 63u) == 0;     // chunk strides keep the alignment
    const size_t chunk_out_bytes = size_t(CHUNK) * 6 * g.np * size_t(elem_bytes(dtype));
    const int64_t chunks = (batch + CHUNK - 1) / CHUNK;
    // a work item = a run of chunks worth >= 256 KB of output: fine enough that a late worker costs little (an item
    // is ~25 us of streaming stores), coarse enough that the ticket is free
    const int64_t run = chunk_out_bytes >= (256u << 10) ? 1 : int64_t((256u << 10) / chunk_out_bytes);
    const int64_t items = (chunks + run - 1) / run;
    const std::function<void(int64_t)> job = [&](int64_t item) {
        const int64_t c0 = item * run, c1 = c0 + run < chunks ? c0 + run : chunks;
        for (int64_t c = c0; c < c1; ++c) {
            const int64_t b0 = c * CHUNK;
            const int boards = int(batch - b0 < CHUNK ? batch - b0 : CHUNK);
            do_chunk(g, rec + size_t(b0) * rec_bytes, boards, dtype, one16, static_cast<uint8_t*>(dense) + size_t(c) * chunk_out_bytes,
                     simd, nt);
        }
        if (nt) _mm_sfence();                               // non-temporal stores: visible before the job is reported done
    };
    int workers = threads;
    if (workers > items) workers = int(items);
    pool().run(workers, items, job);
}

}  // namespace gg
