// Host <-> device movement of PAGEABLE host arrays at PCIe speed.
//
// The reference's hooks hand the engine plain numpy arrays (fuse_np's view slices,
// fusion/_core.py:1579-1587; the destination zarr region of _fuse_chunk_to_zarr,
// :2130-2150).  A cudaMemcpy from pageable memory is staged by the driver through a
// small bounce buffer at a fraction of the link rate.  Here a ring of pinned staging
// buffers is filled (H2D) or drained (D2H) by a pool of worker threads while the DMA
// engine moves the previous piece, so pageable arrays travel at close to the rate of
// pinned ones and the pinning cost is paid once, by the engine.
#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <cstring>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

#include "common.cuh"

namespace mvs {

class CopyPool {
 public:
  explicit CopyPool(int n) {
    for (int i = 0; i < n; ++i) workers_.emplace_back([this] { loop(); });
  }
  ~CopyPool() {
    {
      std::lock_guard<std::mutex> l(m_);
      stop_ = true;
    }
    cv_.notify_all();
    for (auto& t : workers_) t.join();
  }
  int size() const { return (int)workers_.size(); }
  // runs fn(0..n-1) on the pool and the calling thread; returns when all are done
  void parallel_for(int n, const std::function<void(int)>& fn) {
    if (n <= 1 || workers_.empty()) {
      for (int i = 0; i < n; ++i) fn(i);
      return;
    }
    std::atomic<int> next(0), done(0);
    std::mutex dm;
    std::condition_variable dcv;
    auto body = [&] {
      for (;;) {
        int i = next.fetch_add(1);
        if (i >= n) break;
        fn(i);
        if (done.fetch_add(1) + 1 == n) {
          std::lock_guard<std::mutex> l(dm);
          dcv.notify_all();
        }
      }
    };
    const int helpers = std::min<int>(n - 1, (int)workers_.size());
    std::atomic<int> exited(0);
    {
      std::lock_guard<std::mutex> l(m_);
      for (int i = 0; i < helpers; ++i)
        q_.push_back([&] {
          body();
          if (exited.fetch_add(1) + 1 == helpers) {
            std::lock_guard<std::mutex> l2(dm);
            dcv.notify_all();
          }
        });
    }
    cv_.notify_all();
    body();
    std::unique_lock<std::mutex> l(dm);
    dcv.wait(l, [&] { return done.load() >= n && exited.load() >= helpers; });
  }

 private:
  void loop() {
    for (;;) {
      std::function<void()> job;
      {
        std::unique_lock<std::mutex> l(m_);
        cv_.wait(l, [this] { return stop_ || !q_.empty(); });
        if (stop_ && q_.empty()) return;
        job = std::move(q_.front());
        q_.erase(q_.begin());
      }
      job();
    }
  }
  std::vector<std::thread> workers_;
  std::vector<std::function<void()>> q_;
  std::mutex m_;
  std::condition_variable cv_;
  bool stop_ = false;
};

constexpr size_t kPiece = 8u << 20;  // staging piece: 8 MiB
constexpr int kRing = 4;

struct Stager {
  std::mutex mtx;  // one transfer per direction at a time
  void* buf[kRing] = {nullptr, nullptr, nullptr, nullptr};
  cudaEvent_t ev[kRing];
  bool used[kRing] = {false, false, false, false};
  bool ready = false;
  int device = -1;
  int pos = 0;  // next ring slot (persists across calls so small copies pipeline too)
  cudaError_t init() {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (ready && dev == device) return cudaSuccess;
    if (!ready) {
      for (int i = 0; i < kRing; ++i) {
        if ((e = cudaHostAlloc(&buf[i], kPiece, cudaHostAllocPortable)) != cudaSuccess) return e;
      }
    } else {
      for (int i = 0; i < kRing; ++i) cudaEventDestroy(ev[i]);
    }
    for (int i = 0; i < kRing; ++i) {
      if ((e = cudaEventCreateWithFlags(&ev[i], cudaEventDisableTiming)) != cudaSuccess) return e;
      used[i] = false;
    }
    device = dev;
    ready = true;
    return cudaSuccess;
  }
};

static CopyPool& pool() {
  static CopyPool p(std::max(1, std::min(8, (int)std::thread::hardware_concurrency() - 1)));
  return p;
}
static Stager& stager(int dir) {
  static Stager s[2];
  return s[dir];
}

// memcpy of `rows` rows of `width` bytes between pitched layouts, split over the pool
static void pitched_copy(char* dst, size_t dpitch, const char* src, size_t spitch, size_t width,
                         size_t rows) {
  if (dpitch == width && spitch == width) {
    const size_t total = width * rows;
    const int parts = (int)std::min<size_t>(pool().size() + 1, std::max<size_t>(1, total >> 20));
    const size_t per = ((total + parts - 1) / parts + 63) & ~(size_t)63;
    pool().parallel_for(parts, [&](int i) {
      const size_t a = std::min(total, per * i), b = std::min(total, per * (i + 1));
      if (b > a) memcpy(dst + a, src + a, b - a);
    });
    return;
  }
  const int parts = (int)std::min<size_t>(pool().size() + 1, std::max<size_t>(1, (width * rows) >> 20));
  const size_t per = (rows + parts - 1) / parts;
  pool().parallel_for(parts, [&](int i) {
    const size_t a = std::min(rows, per * i), b = std::min(rows, per * (i + 1));
    for (size_t r = a; r < b; ++r) memcpy(dst + r * dpitch, src + r * spitch, width);
  });
}

}  // namespace mvs

using namespace mvs;

extern "C" int mvs_copy_h2d_2d(void* d_dst, size_t d_pitch, const void* h_src, size_t h_pitch,
                               size_t width, size_t rows, void* stream) {
  if (width == 0 || rows == 0) return MVS_OK;
  MVS_REQUIRE(d_dst && h_src, MVS_ERR_INVALID, "NULL pointer");
  MVS_REQUIRE(d_pitch >= width && h_pitch >= width, MVS_ERR_INVALID, "pitch smaller than width");
  MVS_REQUIRE(width <= kPiece, MVS_ERR_UNSUPPORTED, "row of %zu bytes exceeds the staging piece", width);
  Stager& S = stager(0);
  std::lock_guard<std::mutex> lock(S.mtx);
  MVS_CHECK_CUDA(S.init());
  cudaStream_t st = (cudaStream_t)stream;
  const size_t rows_per = std::max<size_t>(1, kPiece / width);
  for (size_t r0 = 0; r0 < rows; r0 += rows_per) {
    const int k = S.pos;
    S.pos = (S.pos + 1) % kRing;
    const size_t n = std::min(rows_per, rows - r0);
    if (S.used[k]) MVS_CHECK_CUDA(cudaEventSynchronize(S.ev[k]));
    pitched_copy((char*)S.buf[k], width, (const char*)h_src + r0 * h_pitch, h_pitch, width, n);
    MVS_CHECK_CUDA(cudaMemcpy2DAsync((char*)d_dst + r0 * d_pitch, d_pitch, S.buf[k], width, width, n,
                                     cudaMemcpyHostToDevice, st));
    MVS_CHECK_CUDA(cudaEventRecord(S.ev[k], st));
    S.used[k] = true;
  }
  return MVS_OK;
}

extern "C" int mvs_copy_d2h_2d(void* h_dst, size_t h_pitch, const void* d_src, size_t d_pitch,
                               size_t width, size_t rows, void* stream) {
  if (width == 0 || rows == 0) return MVS_OK;
  MVS_REQUIRE(h_dst && d_src, MVS_ERR_INVALID, "NULL pointer");
  MVS_REQUIRE(d_pitch >= width && h_pitch >= width, MVS_ERR_INVALID, "pitch smaller than width");
  MVS_REQUIRE(width <= kPiece, MVS_ERR_UNSUPPORTED, "row of %zu bytes exceeds the staging piece", width);
  Stager& S = stager(1);
  std::lock_guard<std::mutex> lock(S.mtx);
  MVS_CHECK_CUDA(S.init());
  cudaStream_t st = (cudaStream_t)stream;
  const size_t rows_per = std::max<size_t>(1, kPiece / width);
  const size_t pieces = (rows + rows_per - 1) / rows_per;
  auto issue = [&](size_t p) -> cudaError_t {
    const int k = (int)(p % kRing);
    const size_t r0 = p * rows_per, n = std::min(rows_per, rows - r0);
    cudaError_t e = cudaMemcpy2DAsync(S.buf[k], width, (const char*)d_src + r0 * d_pitch, d_pitch,
                                      width, n, cudaMemcpyDeviceToHost, st);
    if (e != cudaSuccess) return e;
    return cudaEventRecord(S.ev[k], st);
  };
  // keep kRing - 1 DMA pieces in flight ahead of the CPU drain
  size_t issued = 0;
  for (; issued < std::min<size_t>(pieces, kRing - 1); ++issued) MVS_CHECK_CUDA(issue(issued));
  for (size_t p = 0; p < pieces; ++p) {
    const int k = (int)(p % kRing);
    MVS_CHECK_CUDA(cudaEventSynchronize(S.ev[k]));
    if (issued < pieces) { MVS_CHECK_CUDA(issue(issued)); ++issued; }
    const size_t r0 = p * rows_per, n = std::min(rows_per, rows - r0);
    pitched_copy((char*)h_dst + r0 * h_pitch, h_pitch, (const char*)S.buf[k], width, width, n);
  }
  for (int i = 0; i < kRing; ++i) S.used[i] = false;
  return MVS_OK;
}

// ---------------------------------------------------------------------------------------
// Chunk files of a Zarr directory store (the output side of fuse(output_zarr_url=...),
// ngff_utils.write_sim_to_ome_zarr / _fuse_chunk_to_zarr, fusion/_core.py:1160-1168,
// :2130-2150; input tiles read back the same way): n independent files written from /
// read into host buffers by the copy pool's threads (open + write/read + close each).
// Returns MVS_OK or MVS_ERR_INVALID with the first failing path in the error message.
// ---------------------------------------------------------------------------------------
#include <fcntl.h>
#include <unistd.h>

namespace mvs {
static bool file_rw(const char* path, char* buf, size_t n, bool write) {
  const int fd = write ? open(path, O_WRONLY | O_CREAT | O_TRUNC, 0644) : open(path, O_RDONLY);
  if (fd < 0) return false;
  size_t done = 0;
  while (done < n) {
    const ssize_t r = write ? ::write(fd, buf + done, n - done) : ::read(fd, buf + done, n - done);
    if (r <= 0) break;
    done += (size_t)r;
  }
  close(fd);
  return done == n;
}
}  // namespace mvs

extern "C" int mvs_io_files(const char* const* paths, void* const* bufs, const size_t* sizes, int n,
                            int write) {
  if (n <= 0) return MVS_OK;
  MVS_REQUIRE(paths && bufs && sizes, MVS_ERR_INVALID, "NULL pointer");
  std::atomic<int> failed(-1);
  pool().parallel_for(n, [&](int i) {
    if (!file_rw(paths[i], (char*)bufs[i], sizes[i], write != 0)) {
      int expect = -1;
      failed.compare_exchange_strong(expect, i);
    }
  });
  const int f = failed.load();
  MVS_REQUIRE(f < 0, MVS_ERR_INVALID, "%s failed for %s", write ? "write" : "read", paths[f]);
  return MVS_OK;
}

// ---------------------------------------------------------------------------------------
// Zarr v2 chunk encode / decode on the device.  A chunk of a Zarr array is the C-order
// bytes of a (cz, cy, cx) box, edge chunks padded to the full chunk shape with the fill
// value 0 (the reference states the encoding in ngff_utils.py:372-395, :425-436, and writes
// every chunk -- `write_empty_chunks=True`, `fill_value=0`, :1353-1362).  `chunks_pack`
// gathers a dense (strided) level into chunk-major order -- chunk (iz, iy, ix) at
// ((iz*gy + iy)*gx + ix) * cz*cy*cx elements -- so that every chunk file is ONE contiguous
// device range; `chunks_unpack` is the inverse (input decode).  Pure copies: each dense
// byte is read once and each packed byte written once (HBM bound); lanes run along x, so
// both sides are accessed in runs of cx elements; 16-byte vectors when the chunk rows and
// the dense rows are 16-byte aligned, else element by element.
// ---------------------------------------------------------------------------------------
namespace mvs {

struct ChunkGeom {
  int64_t shape[3], stride[3];  // dense extent / byte strides (x stride == item size)
  int32_t chunk[3], grid[3];
  int es;                       // item size in bytes
};

template <int VEC, bool PACK>  // VEC = bytes per thread along x (16, or the item size)
__global__ void __launch_bounds__(256) chunks_copy_kernel(char* dense, char* packed, ChunkGeom g,
                                                          int64_t n_units) {
  const int64_t row_units = (int64_t)g.chunk[2] * g.es / VEC;  // units per chunk row
  for (int64_t u = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; u < n_units;
       u += (int64_t)gridDim.x * blockDim.x) {
    int64_t r = u / row_units;
    const int64_t xu = u - r * row_units;
    const int cy = (int)(r % g.chunk[1]);
    r /= g.chunk[1];
    const int cz = (int)(r % g.chunk[0]);
    r /= g.chunk[0];
    const int ix = (int)(r % g.grid[2]);
    r /= g.grid[2];
    const int iy = (int)(r % g.grid[1]);
    const int iz = (int)(r / g.grid[1]);
    const int64_t z = (int64_t)iz * g.chunk[0] + cz, y = (int64_t)iy * g.chunk[1] + cy;
    const int64_t xb = ((int64_t)ix * g.chunk[2]) * g.es + xu * VEC;  // byte offset along x
    const int64_t row_bytes = g.shape[2] * g.es;
    const bool in_zy = z < g.shape[0] && y < g.shape[1];
    char* d = dense + z * g.stride[0] + y * g.stride[1] + xb;
    char* p = packed + u * VEC;
    if (VEC == 16) {
      if (PACK) {
        uint4 v = make_uint4(0, 0, 0, 0);
        if (in_zy && xb + 16 <= row_bytes) {
          v = *reinterpret_cast<const uint4*>(d);
        } else if (in_zy && xb < row_bytes) {
          unsigned char t[16];
          for (int b = 0; b < 16; ++b) t[b] = (xb + b < row_bytes) ? (unsigned char)d[b] : 0;
          memcpy(&v, t, 16);
        }
        *reinterpret_cast<uint4*>(p) = v;
      } else if (in_zy && xb < row_bytes) {
        if (xb + 16 <= row_bytes) {
          *reinterpret_cast<uint4*>(d) = *reinterpret_cast<const uint4*>(p);
        } else {
          for (int b = 0; xb + b < row_bytes; ++b) d[b] = p[b];
        }
      }
    } else {
      const bool in = in_zy && xb < row_bytes;
      if (PACK) {
        for (int b = 0; b < VEC; ++b) p[b] = in ? d[b] : 0;
      } else if (in) {
        for (int b = 0; b < VEC; ++b) d[b] = p[b];
      }
    }
  }
}

template <bool PACK>
static int chunks_copy(void* dense, int item_size, const int32_t shape[3], const int64_t stride[3],
                       const int32_t chunk[3], void* packed, cudaStream_t st) {
  MVS_REQUIRE(dense && packed, MVS_ERR_INVALID, "NULL pointer");
  MVS_REQUIRE(item_size == 1 || item_size == 2 || item_size == 4 || item_size == 8, MVS_ERR_UNSUPPORTED,
              "item size %d", item_size);
  ChunkGeom g;
  g.es = item_size;
  int64_t n_bytes = item_size;
  for (int d = 0; d < 3; ++d) {
    MVS_REQUIRE(shape[d] >= 1 && chunk[d] >= 1, MVS_ERR_INVALID, "shape / chunk must be >= 1");
    g.shape[d] = shape[d];
    g.stride[d] = stride[d] * item_size;
    g.chunk[d] = chunk[d];
    g.grid[d] = (shape[d] + chunk[d] - 1) / chunk[d];
    n_bytes *= (int64_t)g.grid[d] * chunk[d];
  }
  MVS_REQUIRE(stride[2] == 1, MVS_ERR_UNSUPPORTED, "the x axis must be contiguous");
  const bool vec = ((int64_t)chunk[2] * item_size) % 16 == 0 && g.stride[0] % 16 == 0 && g.stride[1] % 16 == 0 &&
                   ((uintptr_t)dense % 16) == 0 && ((uintptr_t)packed % 16) == 0;
  const int64_t units = n_bytes / (vec ? 16 : item_size);
  const int blocks = (int)std::min<int64_t>((units + 255) / 256, 148 * 32);
  if (vec) {
    chunks_copy_kernel<16, PACK><<<blocks, 256, 0, st>>>((char*)dense, (char*)packed, g, units);
  } else if (item_size == 1) {
    chunks_copy_kernel<1, PACK><<<blocks, 256, 0, st>>>((char*)dense, (char*)packed, g, units);
  } else if (item_size == 2) {
    chunks_copy_kernel<2, PACK><<<blocks, 256, 0, st>>>((char*)dense, (char*)packed, g, units);
  } else if (item_size == 4) {
    chunks_copy_kernel<4, PACK><<<blocks, 256, 0, st>>>((char*)dense, (char*)packed, g, units);
  } else {
    chunks_copy_kernel<8, PACK><<<blocks, 256, 0, st>>>((char*)dense, (char*)packed, g, units);
  }
  MVS_CHECK_CUDA(cudaGetLastError());
  return MVS_OK;
}

// per-thread staging for the chunk store / load pipeline: one pinned buffer (grow-only) and
// one stream per pool thread, so the DMA of one chunk overlaps the file I/O of the others
struct ChunkLane {
  void* pinned = nullptr;
  size_t cap = 0;
  cudaStream_t st = nullptr;
  int device = -1;
  cudaError_t prepare(int dev, size_t bytes) {
    cudaError_t e = cudaSetDevice(dev);
    if (e != cudaSuccess) return e;
    if (st == nullptr || device != dev) {
      if (st) cudaStreamDestroy(st);
      if ((e = cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking)) != cudaSuccess) return e;
      device = dev;
    }
    if (cap < bytes) {
      if (pinned) cudaFreeHost(pinned);
      pinned = nullptr;
      cap = 0;
      if ((e = cudaHostAlloc(&pinned, bytes, cudaHostAllocPortable)) != cudaSuccess) return e;
      cap = bytes;
    }
    return cudaSuccess;
  }
};

static thread_local ChunkLane t_lane;

}  // namespace mvs

extern "C" int mvs_chunks_pack(const void* d_dense, int item_size, const int32_t shape[3],
                               const int64_t stride[3], const int32_t chunk[3], void* d_packed,
                               void* stream) {
  return chunks_copy<true>(const_cast<void*>(d_dense), item_size, shape, stride, chunk, d_packed,
                           (cudaStream_t)stream);
}

extern "C" int mvs_chunks_unpack(const void* d_packed, int item_size, const int32_t shape[3],
                                 const int64_t stride[3], const int32_t chunk[3], void* d_dense,
                                 void* stream) {
  return chunks_copy<false>(d_dense, item_size, shape, stride, chunk, const_cast<void*>(d_packed),
                            (cudaStream_t)stream);
}

extern "C" int mvs_chunks_store(const void* d_packed, size_t chunk_bytes, int n,
                                const char* const* paths, void* stream) {
  if (n <= 0) return MVS_OK;
  MVS_REQUIRE(d_packed && paths && chunk_bytes > 0, MVS_ERR_INVALID, "NULL pointer / empty chunk");
  int dev = 0;
  MVS_CHECK_CUDA(cudaGetDevice(&dev));
  cudaEvent_t ready;
  MVS_CHECK_CUDA(cudaEventCreateWithFlags(&ready, cudaEventDisableTiming));
  MVS_CHECK_CUDA(cudaEventRecord(ready, (cudaStream_t)stream));
  std::atomic<int> failed(-1), cuda_err(0);
  pool().parallel_for(n, [&](int i) {
    ChunkLane& L = t_lane;
    cudaError_t e = L.prepare(dev, chunk_bytes);
    if (e == cudaSuccess) e = cudaStreamWaitEvent(L.st, ready, 0);
    if (e == cudaSuccess)
      e = cudaMemcpyAsync(L.pinned, (const char*)d_packed + (size_t)i * chunk_bytes, chunk_bytes,
                          cudaMemcpyDeviceToHost, L.st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(L.st);
    if (e != cudaSuccess) {
      cuda_err.store((int)e);
      return;
    }
    if (!file_rw(paths[i], (char*)L.pinned, chunk_bytes, true)) {
      int expect = -1;
      failed.compare_exchange_strong(expect, i);
    }
  });
  cudaEventDestroy(ready);
  MVS_REQUIRE(cuda_err.load() == 0, MVS_ERR_CUDA, "chunk download failed: %s",
              cudaGetErrorString((cudaError_t)cuda_err.load()));
  const int f = failed.load();
  MVS_REQUIRE(f < 0, MVS_ERR_INVALID, "write failed for %s", paths[f]);
  return MVS_OK;
}

extern "C" int mvs_chunks_load(void* d_packed, size_t chunk_bytes, int n, const char* const* paths,
                               void* stream) {
  if (n <= 0) return MVS_OK;
  MVS_REQUIRE(d_packed && paths && chunk_bytes > 0, MVS_ERR_INVALID, "NULL pointer / empty chunk");
  int dev = 0;
  MVS_CHECK_CUDA(cudaGetDevice(&dev));
  cudaEvent_t ready;
  MVS_CHECK_CUDA(cudaEventCreateWithFlags(&ready, cudaEventDisableTiming));
  MVS_CHECK_CUDA(cudaEventRecord(ready, (cudaStream_t)stream));
  std::atomic<int> failed(-1), cuda_err(0);
  pool().parallel_for(n, [&](int i) {
    ChunkLane& L = t_lane;
    cudaError_t e = L.prepare(dev, chunk_bytes);
    if (e == cudaSuccess) e = cudaStreamWaitEvent(L.st, ready, 0);
    if (e == cudaSuccess) {
      char* dst = (char*)d_packed + (size_t)i * chunk_bytes;
      if (access(paths[i], F_OK) != 0) {
        e = cudaMemsetAsync(dst, 0, chunk_bytes, L.st);  // a missing chunk reads as the fill value
      } else if (file_rw(paths[i], (char*)L.pinned, chunk_bytes, false)) {
        e = cudaMemcpyAsync(dst, L.pinned, chunk_bytes, cudaMemcpyHostToDevice, L.st);
      } else {
        int expect = -1;
        failed.compare_exchange_strong(expect, i);
      }
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(L.st);
    if (e != cudaSuccess) cuda_err.store((int)e);
  });
  cudaEventDestroy(ready);
  MVS_REQUIRE(cuda_err.load() == 0, MVS_ERR_CUDA, "chunk upload failed: %s",
              cudaGetErrorString((cudaError_t)cuda_err.load()));
  const int f = failed.load();
  MVS_REQUIRE(f < 0, MVS_ERR_INVALID, "read failed for %s (short or unreadable chunk file)", paths[f]);
  return MVS_OK;
}
