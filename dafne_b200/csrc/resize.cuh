// Pillow-exact bilinear resize of uint8 planes (see resize.cu).
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

namespace dafne {
size_t resize_tmp_bytes(int planes, int H, int nw);
// in [planes][H][W] uint8 -> out [planes][nh][nw] uint8; tmp: resize_tmp_bytes(planes, H, nw) bytes, needed when both
// sizes change.
int launch_resize_bilinear_u8(const uint8_t* in, int planes, int H, int W, uint8_t* out, int nh, int nw, uint8_t* tmp,
                              cudaStream_t stream);
}  // namespace dafne
