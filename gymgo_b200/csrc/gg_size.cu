// gg_size.cu - compiled once per board size with -DGG_N=<n>; instantiates every kernel for that size and
// exports its launch table.  Splitting by size keeps each nvcc job small and lets the build run them in
// parallel.
#include "gg_kernels.cuh"
#include "gg_sliced.cuh"

#ifndef GG_N
#error "compile with -DGG_N=<board size>"
#endif

#define GG_CAT2(a, b) a##b
#define GG_CAT(a, b) GG_CAT2(a, b)

namespace gg {
template <>
bool Launch<Geo<GG_N>>::launch_sliced(const RolloutArgs& a, cudaStream_t s) {
    typedef Geo<GG_N> G;
    switch (a.slice_k) {
        case 2: return LaunchSliced<G, 2>::go(a, s);
        case 3: return LaunchSliced<G, 3>::go(a, s);
        case 4: return LaunchSliced<G, 4>::go(a, s);
    }
    return false;
}

extern const SizeVTable GG_CAT(vtable_n, GG_N);
const SizeVTable GG_CAT(vtable_n, GG_N) = Launch<Geo<GG_N>>::table();
}  // namespace gg
