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
