// gg_size.cu - compiled once per board size with -DGG_N=<n>; instantiates every kernel for that size and
// exports its launch table.  Splitting by size keeps each nvcc job small and lets the build run them in
// parallel.
#include "gg_kernels.cuh"

#ifndef GG_N
#error "compile with -DGG_N=<board size>"
#endif

#define GG_CAT2(a, b) a##b
#define GG_CAT(a, b) GG_CAT2(a, b)

namespace gg {
extern const SizeVTable GG_CAT(vtable_n, GG_N);
const SizeVTable GG_CAT(vtable_n, GG_N) = Launch<Geo<GG_N>>::table();
}  // namespace gg
