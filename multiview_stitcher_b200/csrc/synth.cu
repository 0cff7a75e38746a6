// Synthetic tile generator (benchmark / test inputs).
//
// Integer-only value noise sampled at integer global coordinates, so tiles cut
// from the same ground truth agree exactly in their overlaps and the same
// values can be regenerated anywhere (host mirror: synthetic.py ground_truth).
//   coarse octave: lattice 16 px, amplitude 4096
//   fine octave  : lattice  4 px, amplitude 1024
//   voxel noise  : amplitude 256
#include "common.cuh"

namespace mvs {

__host__ __device__ inline uint32_t lattice_octave(uint32_t seed, int64_t z, int64_t y,
                                                   int64_t x, int shift, uint32_t amp_mask) {
  const int64_t cell = (int64_t)1 << shift;
  // floor division for negative coordinates
  int64_t cz = z >> shift, cy = y >> shift, cx = x >> shift;
  uint32_t wz = (uint32_t)(z - (cz << shift)), wy = (uint32_t)(y - (cy << shift)),
           wx = (uint32_t)(x - (cx << shift));
  uint32_t c = (uint32_t)cell;
  uint64_t acc = 0;
  for (int dz = 0; dz < 2; ++dz)
    for (int dy = 0; dy < 2; ++dy)
      for (int dx = 0; dx < 2; ++dx) {
        uint32_t h = hash3(seed, cz + dz, cy + dy, cx + dx) & amp_mask;
        uint32_t w = (dz ? wz : c - wz) * (dy ? wy : c - wy) * (dx ? wx : c - wx);
        acc += (uint64_t)h * w;
      }
  return (uint32_t)(acc >> (3 * shift));
}

__host__ __device__ inline uint32_t ground_truth(uint32_t seed, int64_t z, int64_t y,
                                                 int64_t x) {
  uint32_t v = lattice_octave(seed, z, y, x, 4, 4095u);
  v += lattice_octave(seed + 1u, z, y, x, 2, 1023u);
  v += hash3(seed + 2u, z, y, x) & 255u;
  return v;  // < 5376
}

__global__ void synth_kernel(void* out, int dtype, int nz, int ny, int nx, int64_t sz,
                             int64_t sy, int64_t sx, int64_t oz, int64_t oy, int64_t ox,
                             uint32_t seed) {
  const int64_t n = (int64_t)nz * ny * nx;
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t step = (int64_t)gridDim.x * blockDim.x;
  for (; i < n; i += step) {
    int x = (int)(i % nx);
    int y = (int)((i / nx) % ny);
    int z = (int)(i / ((int64_t)nx * ny));
    uint32_t v = ground_truth(seed, oz + z, oy + y, ox + x);
    int64_t o = z * sz + y * sy + x * sx;
    if (dtype == MVS_F32)
      reinterpret_cast<float*>(out)[o] = (float)v * (1.0f / 8192.0f);
    else if (dtype == MVS_U16)
      reinterpret_cast<unsigned short*>(out)[o] = (unsigned short)v;
    else
      reinterpret_cast<unsigned char*>(out)[o] = (unsigned char)(v >> 5);
  }
}

}  // namespace mvs

extern "C" int mvs_synth_tile(void* d_out, int dtype, const int32_t shape[3],
                              const int64_t stride[3], const int64_t origin[3], uint32_t seed,
                              void* stream) {
  using namespace mvs;
  MVS_REQUIRE(d_out && shape && stride && origin, MVS_ERR_INVALID, "NULL pointer");
  MVS_REQUIRE(dtype >= MVS_U8 && dtype <= MVS_F32, MVS_ERR_INVALID, "bad dtype %d", dtype);
  const int64_t n = (int64_t)shape[0] * shape[1] * shape[2];
  if (n <= 0) return MVS_OK;
  int blocks = (int)((n + 255) / 256 < 148 * 32 ? (n + 255) / 256 : 148 * 32);
  synth_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(d_out, dtype, shape[0], shape[1],
                                                        shape[2], stride[0], stride[1],
                                                        stride[2], origin[0], origin[1],
                                                        origin[2], seed);
  MVS_CHECK_CUDA(cudaGetLastError());
  return MVS_OK;
}
