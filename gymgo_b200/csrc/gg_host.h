// gg_host.h - internal interface of the host codec (gg_host.cpp, compiled by the host compiler) used by gg_api.cu
#pragma once
#include <stdint.h>

namespace gg {
enum { HOST_U8 = 0, HOST_F32 = 1, HOST_F64 = 2, HOST_BF16 = 3, HOST_F16 = 4 };   // == GG_U8 .. GG_F16
// packed records [batch] in host memory -> dense [batch,6,n,n] of dtype in host memory on `threads` threads
void host_unpack(const uint8_t* rec, int64_t batch, int n, int lpb, int rpl, int wordbits, int rec_bytes, int dtype,
                 void* dense, int threads);
const char* host_unpack_path();
}  // namespace gg
