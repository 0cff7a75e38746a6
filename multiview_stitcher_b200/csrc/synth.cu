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


// ---------------------------------------------------------------------------
// Band-limited analytic ground truth: a sum of K plane waves
//   f(p) = base + sum_k a_k sin(2 pi w_k . p + phi_k)
// evaluable at ANY real position p, so tiles may sit at fractional origins
// (sub-pixel jitter) and still agree in their overlaps up to their own sampling.
// The wave of term k at voxel (z, y, x) of a tile factorises into per-axis
// complex exponentials, tabulated on the host in float64 (exact for large
// origins) and handed over as float2 tables e_axis[k * n_axis + i]:
//   sin(theta) = Im( c_k * Ez[k,z] * Ey[k,y] * Ex[k,x] ),  c_k = a_k e^{i phi_k}.
// A block owns 256 columns x ROWS rows of one plane; the row coefficients
// c_k Ez Ey live in shared memory, Ex[k, x] is loaded once per term and thread.
// Per-tile noise (independent between tiles like camera noise) is a counter hash.
constexpr int kFieldRows = 16;
constexpr int kFieldMaxTerms = 128;

__global__ void __launch_bounds__(256)
synth_field_kernel(void* out, int dtype, int nz, int ny, int nx, int64_t sz, int64_t sy,
                   const float2* __restrict__ ez, const float2* __restrict__ ey,
                   const float2* __restrict__ ex, const float2* __restrict__ coef, int K,
                   float base, float out_scale, float noise_amp, uint32_t noise_seed) {
  __shared__ float2 rowc[kFieldRows][kFieldMaxTerms];
  const int z = blockIdx.z;
  const int y0 = blockIdx.y * kFieldRows;
  const int x = blockIdx.x * 256 + threadIdx.x;
  for (int i = threadIdx.x; i < kFieldRows * K; i += 256) {
    const int r = i / K, k = i - r * K;
    const int y = min(y0 + r, ny - 1);
    const float2 c = coef[k], a = ez[(int64_t)k * nz + z], b = ey[(int64_t)k * ny + y];
    const float2 ab = make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
    rowc[r][k] = make_float2(c.x * ab.x - c.y * ab.y, c.x * ab.y + c.y * ab.x);
  }
  __syncthreads();
  if (x >= nx) return;
  float acc[kFieldRows];
#pragma unroll
  for (int r = 0; r < kFieldRows; ++r) acc[r] = 0.f;
  for (int k = 0; k < K; ++k) {
    const float2 e = __ldg(ex + (int64_t)k * nx + x);
#pragma unroll
    for (int r = 0; r < kFieldRows; ++r) {
      const float2 c = rowc[r][k];
      acc[r] = fmaf(c.x, e.y, fmaf(c.y, e.x, acc[r]));  // Im(c * e)
    }
  }
#pragma unroll
  for (int r = 0; r < kFieldRows; ++r) {
    const int y = y0 + r;
    if (y >= ny) break;
    float v = base + acc[r];
    if (noise_amp != 0.f) {
      const uint32_t h = hash3(noise_seed, z, y, x);
      v += noise_amp * ((float)(h >> 8) * (1.0f / 16777216.0f) - 0.5f);
    }
    v = fminf(fmaxf(v, 0.f), 0.99999994f);
    const int64_t o = (int64_t)z * sz + (int64_t)y * sy + x;
    if (dtype == MVS_F32)
      reinterpret_cast<float*>(out)[o] = v * out_scale;
    else if (dtype == MVS_U16)
      reinterpret_cast<unsigned short*>(out)[o] = (unsigned short)(v * out_scale);
    else
      reinterpret_cast<unsigned char*>(out)[o] = (unsigned char)(v * out_scale);
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

extern "C" int mvs_synth_field(void* d_out, int dtype, const int32_t shape[3],
                               const int64_t stride[3], const float* d_ez, const float* d_ey,
                               const float* d_ex, const float* d_coef, int n_terms, float base,
                               float out_scale, float noise_amp, uint32_t noise_seed,
                               void* stream) {
  using namespace mvs;
  MVS_REQUIRE(d_out && shape && stride && d_ez && d_ey && d_ex && d_coef, MVS_ERR_INVALID,
              "NULL pointer");
  MVS_REQUIRE(dtype >= MVS_U8 && dtype <= MVS_F32, MVS_ERR_INVALID, "bad dtype %d", dtype);
  MVS_REQUIRE(n_terms >= 1 && n_terms <= kFieldMaxTerms, MVS_ERR_INVALID,
              "n_terms %d outside 1..%d", n_terms, kFieldMaxTerms);
  MVS_REQUIRE(stride[2] == 1, MVS_ERR_UNSUPPORTED, "tile rows must be contiguous");
  if ((int64_t)shape[0] * shape[1] * shape[2] <= 0) return MVS_OK;
  MVS_REQUIRE(shape[0] <= 65535, MVS_ERR_UNSUPPORTED, "z extent %d > 65535", shape[0]);
  dim3 grid((unsigned)((shape[2] + 255) / 256), (unsigned)((shape[1] + kFieldRows - 1) / kFieldRows),
            (unsigned)shape[0]);
  synth_field_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(
      d_out, dtype, shape[0], shape[1], shape[2], stride[0], stride[1],
      reinterpret_cast<const float2*>(d_ez), reinterpret_cast<const float2*>(d_ey),
      reinterpret_cast<const float2*>(d_ex), reinterpret_cast<const float2*>(d_coef), n_terms,
      base, out_scale, noise_amp, noise_seed);
  MVS_CHECK_CUDA(cudaGetLastError());
  return MVS_OK;
}
