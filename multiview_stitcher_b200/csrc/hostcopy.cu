// Host <-> device movement of PAGEABLE host arrays at PCIe speed.
//
// The reference's hooks hand the engine plain numpy arrays (fuse_np's view slices,
// fusion/_core.py:1579-1587; the destination zarr region of _fuse_chunk_to_zarr,
// :2130-2150).  A cudaMemcpy from pageable memory is staged by the driver through a
// small bounce buffer at a fraction of the link rate.  Here a transfer is cut into 1 MiB
// pieces that travel through a ring of pinned slots: a pool of threads copies between the
// user's array and the slots (several pieces at a time; downloads with cache-bypassing
// stores) while the calling thread alone enqueues the DMAs and waits for their events, so
// pageable arrays travel at close to the rate of pinned ones, uploads and downloads overlap,
// and the pinning cost is paid once, by the engine.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>
#if defined(__SSE2__)
#include <emmintrin.h>
#endif

#include "common.cuh"

namespace mvs {

static inline void cpu_relax() {
#if defined(__SSE2__)
  _mm_pause();
#else
  std::this_thread::yield();
#endif
}

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
  // fire-and-forget job (the caller tracks completion itself)
  void submit(std::function<void()> job) {
    {
      std::lock_guard<std::mutex> l(m_);
      q_.push_back(std::move(job));
      pending_.fetch_add(1, std::memory_order_release);
    }
    if (sleepers_.load(std::memory_order_acquire) > 0) cv_.notify_one();
  }
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
      pending_.fetch_add(helpers, std::memory_order_release);
    }
    cv_.notify_all();
    body();
    std::unique_lock<std::mutex> l(dm);
    dcv.wait(l, [&] { return done.load() >= n && exited.load() >= helpers; });
  }

 private:
  // A worker that has just run a job keeps polling for the next one for a while before it goes
  // to sleep: the pieces of a transfer arrive every few tens of microseconds, and waking a
  // sleeping thread costs more than copying a piece (measured on the B200 box: a transfer whose
  // workers sleep between pieces runs at a third of the rate).
  void loop() {
    for (;;) {
      std::function<void()> job;
      const auto t0 = std::chrono::steady_clock::now();
      for (;;) {
        if (pending_.load(std::memory_order_acquire) > 0) {
          std::lock_guard<std::mutex> l(m_);
          if (!q_.empty()) {
            job = std::move(q_.front());
            q_.erase(q_.begin());
            pending_.fetch_sub(1, std::memory_order_relaxed);
            break;
          }
        }
        if (std::chrono::steady_clock::now() - t0 > std::chrono::microseconds(400)) {
          std::unique_lock<std::mutex> l(m_);
          sleepers_.fetch_add(1);
          cv_.wait(l, [this] { return stop_ || !q_.empty(); });
          sleepers_.fetch_sub(1);
          if (stop_ && q_.empty()) return;
          job = std::move(q_.front());
          q_.erase(q_.begin());
          pending_.fetch_sub(1, std::memory_order_relaxed);
          break;
        }
        cpu_relax();
      }
      job();
    }
  }
  std::atomic<int> pending_{0}, sleepers_{0};
  std::vector<std::thread> workers_;
  std::vector<std::function<void()>> q_;
  std::mutex m_;
  std::condition_variable cv_;
  bool stop_ = false;
};

// staging piece (default 1 MiB; MVS_COPY_PIECE_KB overrides it for experiments).  Measured on hook C's C2 step:
// 256 KiB 28.7 ms, 512 KiB 19.5, 768 KiB 16.4, 1 MiB 15.1, 1.5 MiB 15.7, 2 MiB 16.1, 4 MiB 16.1
static size_t piece_bytes() {
  static const size_t v = [] {
    const char* e = getenv("MVS_COPY_PIECE_KB");
    const long kb = e ? atol(e) : 0;
    return kb >= 64 && kb <= (64 << 10) ? (size_t)kb << 10 : (size_t)1 << 20;
  }();
  return v;
}
#define kPiece (piece_bytes())
static bool stream_stores_enabled() {
  static const bool v = [] {
    const char* e = getenv("MVS_COPY_NT");
    return !(e && e[0] == '0');
  }();
  return v;
}

static CopyPool& pool() {
  static CopyPool p([] {
    const char* e = getenv("MVS_COPY_THREADS");
    const int want = e ? atoi(e) : 0;
    // one process per GPU shares the host's cores with its node-local peers (torchrun exports
    // LOCAL_WORLD_SIZE): polling workers must not oversubscribe them
    const char* lw = getenv("LOCAL_WORLD_SIZE");
    const int peers = std::max(1, lw ? atoi(lw) : 1);
    const int hw = std::max(1, (int)std::thread::hardware_concurrency() / peers);
    return want > 0 ? std::min(want, 64) : std::max(2, std::min(8, hw - 1));
  }());
  return p;
}

// memcpy whose stores bypass the cache (the destination of a download is written once and
// not read by this thread again: no read-for-ownership traffic on the host memory bus)
static void copy_stream_stores(char* dst, const char* src, size_t n) {
#if defined(__SSE2__)
  if (n >= 256 && stream_stores_enabled()) {
    const size_t head = (16 - ((uintptr_t)dst & 15)) & 15;
    if (head) {
      memcpy(dst, src, head);
      dst += head, src += head, n -= head;
    }
    const size_t body = n & ~(size_t)63;
    for (size_t i = 0; i < body; i += 64) {
      const __m128i a = _mm_loadu_si128((const __m128i*)(src + i));
      const __m128i b = _mm_loadu_si128((const __m128i*)(src + i + 16));
      const __m128i c = _mm_loadu_si128((const __m128i*)(src + i + 32));
      const __m128i d = _mm_loadu_si128((const __m128i*)(src + i + 48));
      _mm_stream_si128((__m128i*)(dst + i), a);
      _mm_stream_si128((__m128i*)(dst + i + 16), b);
      _mm_stream_si128((__m128i*)(dst + i + 32), c);
      _mm_stream_si128((__m128i*)(dst + i + 48), d);
    }
    dst += body, src += body, n -= body;
  }
#endif
  if (n) memcpy(dst, src, n);
}

// One staging ring per direction: kSlots pinned pieces.  ONE thread -- the caller -- talks to
// the CUDA driver (DMA enqueue, event waits); the pool's threads only run the host copies
// between the user's pageable array and the pinned pieces, several pieces at a time, and
// report back through a flag per slot.  (Measured on the B200 box: letting every worker
// enqueue its own DMAs and poll its own events makes an upload and a download that run side
// by side 8x slower than either alone -- the driver serialises the calls.)
#ifndef MVS_COPY_SLOTS
#define MVS_COPY_SLOTS 16
#define MVS_COPY_FILL 12
#define MVS_COPY_DMA 6
#endif
constexpr int kSlots = MVS_COPY_SLOTS;
constexpr int kFillAhead = MVS_COPY_FILL;  // upload: pieces being filled by the pool (the other slots hold DMAs in flight)
constexpr int kDmaAhead = MVS_COPY_DMA;    // download: DMAs in flight (the other slots are being drained by the pool)

struct Ring {
  std::mutex mtx;  // one transfer per direction at a time
  void* buf[kSlots] = {};
  size_t cap = 0;
  cudaEvent_t ev[kSlots] = {};
  bool dma_pending[kSlots] = {};
  size_t pos = 0;  // first slot of the next transfer: consecutive calls keep rotating through the ring
  int device = -1;
  // host-copy completion, signalled by pool threads (the caller polls: see CopyPool::loop)
  std::atomic<int> host_done[kSlots] = {};

  cudaError_t prepare(int dev, size_t bytes) {
    cudaError_t e;
    if (device != dev) {
      for (int i = 0; i < kSlots; ++i) {
        if (dma_pending[i]) cudaEventSynchronize(ev[i]);
        if (ev[i]) cudaEventDestroy(ev[i]);
        if ((e = cudaEventCreateWithFlags(&ev[i], cudaEventDisableTiming)) != cudaSuccess) return e;
        dma_pending[i] = false;
      }
      device = dev;
    }
    if (cap < bytes) {
      for (int i = 0; i < kSlots; ++i) {
        if (dma_pending[i]) cudaEventSynchronize(ev[i]);
        dma_pending[i] = false;
        if (buf[i]) cudaFreeHost(buf[i]);
        buf[i] = nullptr;
      }
      cap = 0;
      for (int i = 0; i < kSlots; ++i)
        if ((e = cudaHostAlloc(&buf[i], bytes + 4096, cudaHostAllocPortable)) != cudaSuccess) return e;
      cap = bytes;
    }
    return cudaSuccess;
  }
  // The piece's place inside slot k: half a page away (mod 4 KiB) from the user's address, so
  // that the loads and stores of the host copy never share their low 12 address bits (when they
  // do -- e.g. 2 MiB pieces of a page-aligned array -- every load falsely depends on the store
  // before it and the copy runs at half speed; measured on the B200 box).
  char* at(int k, const void* user) const {
    const uintptr_t want = (((uintptr_t)user & 4095) + 2048) & 4095 & ~(uintptr_t)63;
    return (char*)buf[k] + want;
  }
  void mark(int slot) { host_done[slot].store(1, std::memory_order_release); }
  void wait_host(int slot) {
    for (int spins = 0; host_done[slot].load(std::memory_order_acquire) == 0; ++spins) {
      if (spins < (1 << 16)) cpu_relax();
      else std::this_thread::yield();
    }
  }
  void clear(int slot) { host_done[slot].store(0, std::memory_order_relaxed); }
};

static Ring& ring(int dir) {
  static Ring r[2];
  return r[dir];
}

// a transfer of `planes` planes of `rows` rows of `width` bytes cut into pieces: byte ranges
// when both sides are contiguous, else runs of whole rows inside one plane
struct Pieces {
  bool flat;
  size_t width, rows, planes, unit, per_plane, n, buf_bytes;
  Pieces(size_t width_, size_t rows_, size_t planes_, size_t d_pitch, size_t d_plane, size_t h_pitch,
         size_t h_plane)
      : width(width_), rows(rows_), planes(planes_) {
    flat = d_pitch == width && h_pitch == width &&
           (planes == 1 || (d_plane == width * rows && h_plane == width * rows));
    // small transfers are cut finer (>= 256 KiB) so that all pool threads work on them
    const size_t total = width * rows * planes;
    const size_t piece = std::min<size_t>(kPiece, std::max<size_t>((size_t)256 << 10, ((total / 16) + 65535) & ~(size_t)65535));
    if (flat) {
      unit = piece;
      per_plane = 0;
      n = (total + unit - 1) / unit;
      buf_bytes = kPiece;
    } else {
      unit = std::max<size_t>(1, piece / width);  // rows per piece
      per_plane = (rows + unit - 1) / unit;
      n = per_plane * planes;
      buf_bytes = std::max(kPiece, width);
    }
  }
  // piece p: byte offsets on both sides and its extent (flat: n_rows == 0 and `bytes` valid)
  void locate(size_t p, size_t d_pitch, size_t d_plane, size_t h_pitch, size_t h_plane, size_t* d_off,
              size_t* h_off, size_t* n_rows, size_t* bytes) const {
    if (flat) {
      *d_off = *h_off = p * unit;
      *n_rows = 0;
      *bytes = std::min(unit, width * rows * planes - p * unit);
      return;
    }
    const size_t pl = p / per_plane, r0 = (p % per_plane) * unit;
    *d_off = pl * d_plane + r0 * d_pitch;
    *h_off = pl * h_plane + r0 * h_pitch;
    *n_rows = std::min(unit, rows - r0);
    *bytes = *n_rows * width;
  }
};

// memcpy of `rows` rows of `width` bytes between pitched layouts
template <bool STREAM>
static void rows_copy(char* dst, size_t dpitch, const char* src, size_t spitch, size_t width, size_t rows) {
  if (dpitch == width && spitch == width) {
    if (STREAM) copy_stream_stores(dst, src, width * rows);
    else memcpy(dst, src, width * rows);
    return;
  }
  for (size_t r = 0; r < rows; ++r) {
    if (STREAM) copy_stream_stores(dst + r * dpitch, src + r * spitch, width);
    else memcpy(dst + r * dpitch, src + r * spitch, width);
  }
}

static int staged_upload(char* d_dst, size_t d_pitch, size_t d_plane, const char* h_src, size_t h_pitch,
                         size_t h_plane, size_t width, size_t rows, size_t planes, cudaStream_t st) {
  if (width == 0 || rows == 0 || planes == 0) return MVS_OK;
  MVS_REQUIRE(d_dst && h_src, MVS_ERR_INVALID, "NULL pointer");
  MVS_REQUIRE(d_pitch >= width && h_pitch >= width, MVS_ERR_INVALID, "pitch smaller than width");
  int dev = 0;
  MVS_CHECK_CUDA(cudaGetDevice(&dev));
  const Pieces P(width, rows, planes, d_pitch, d_plane, h_pitch, h_plane);
  Ring& R = ring(0);
  std::lock_guard<std::mutex> lock(R.mtx);
  MVS_CHECK_CUDA(R.prepare(dev, P.buf_bytes));
  // pieces are handed to the pool as slots free up (a slot is free once the DMA that last read it
  // has completed) and their DMAs are enqueued on `stream`, in order, as the fills complete
  size_t dispatched = 0, issued = 0;
  const size_t base = R.pos;
  R.pos = (R.pos + P.n) % kSlots;
  cudaError_t e = cudaSuccess;
  while (issued < P.n && e == cudaSuccess) {
    while (dispatched < P.n && dispatched < issued + kFillAhead) {
      const int k = (int)((base + dispatched) % kSlots);
      if (R.dma_pending[k]) {
        if ((e = cudaEventSynchronize(R.ev[k])) != cudaSuccess) break;
        R.dma_pending[k] = false;
      }
      R.clear(k);
      const size_t p = dispatched++;
      pool().submit([&R, &P, k, p, h_src, d_pitch, d_plane, h_pitch, h_plane, width] {
        size_t d_off, h_off, n_rows, bytes;
        P.locate(p, d_pitch, d_plane, h_pitch, h_plane, &d_off, &h_off, &n_rows, &bytes);
        char* stage = R.at(k, h_src + h_off);
        if (n_rows == 0) memcpy(stage, h_src + h_off, bytes);
        else rows_copy<false>(stage, width, h_src + h_off, h_pitch, width, n_rows);
        R.mark(k);
      });
    }
    if (e != cudaSuccess) break;
    const int k = (int)((base + issued) % kSlots);
    R.wait_host(k);
    size_t d_off, h_off, n_rows, bytes;
    P.locate(issued, d_pitch, d_plane, h_pitch, h_plane, &d_off, &h_off, &n_rows, &bytes);
    const char* stage = R.at(k, h_src + h_off);
    if (n_rows == 0) e = cudaMemcpyAsync(d_dst + d_off, stage, bytes, cudaMemcpyHostToDevice, st);
    else e = cudaMemcpy2DAsync(d_dst + d_off, d_pitch, stage, width, width, n_rows, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) e = cudaEventRecord(R.ev[k], st);
    R.dma_pending[k] = true;
    ++issued;
  }
  for (size_t p = issued; p < dispatched; ++p) R.wait_host((int)((base + p) % kSlots));  // error path: no job may outlive the call
  MVS_REQUIRE(e == cudaSuccess, MVS_ERR_CUDA, "staged upload failed: %s", cudaGetErrorString(e));
  return MVS_OK;
}

// Many contiguous host arrays -> device buffers as ONE pipelined transfer (the crops of a batch of
// pairs: per-array calls leave the pool idle between arrays; measured 2 x 40 crops of 2.5 MB:
// ~10 ms array by array).
static int staged_upload_many(int n, char* const* d_dst, const char* const* h_src, const size_t* bytes,
                              cudaStream_t st) {
  struct Seg { char* d; const char* h; size_t n; };
  std::vector<Seg> segs;
  size_t total = 0;
  for (int i = 0; i < n; ++i) total += bytes[i];
  if (total == 0) return MVS_OK;
  const size_t piece = std::min<size_t>(kPiece, std::max<size_t>((size_t)256 << 10, ((total / 64) + 65535) & ~(size_t)65535));
  for (int i = 0; i < n; ++i) {
    MVS_REQUIRE(bytes[i] == 0 || (d_dst[i] && h_src[i]), MVS_ERR_INVALID, "array %d: NULL pointer", i);
    for (size_t a = 0; a < bytes[i]; a += piece) segs.push_back({d_dst[i] + a, h_src[i] + a, std::min(piece, bytes[i] - a)});
  }
  int dev = 0;
  MVS_CHECK_CUDA(cudaGetDevice(&dev));
  Ring& R = ring(0);
  std::lock_guard<std::mutex> lock(R.mtx);
  MVS_CHECK_CUDA(R.prepare(dev, kPiece));
  const size_t N = segs.size(), base = R.pos;
  R.pos = (R.pos + N) % kSlots;
  size_t dispatched = 0, issued = 0;
  cudaError_t e = cudaSuccess;
  const Seg* S = segs.data();
  while (issued < N && e == cudaSuccess) {
    while (dispatched < N && dispatched < issued + kFillAhead) {
      const int k = (int)((base + dispatched) % kSlots);
      if (R.dma_pending[k]) {
        if ((e = cudaEventSynchronize(R.ev[k])) != cudaSuccess) break;
        R.dma_pending[k] = false;
      }
      R.clear(k);
      const size_t p = dispatched++;
      pool().submit([&R, S, k, p] {
        memcpy(R.at(k, S[p].h), S[p].h, S[p].n);
        R.mark(k);
      });
    }
    if (e != cudaSuccess) break;
    const int k = (int)((base + issued) % kSlots);
    R.wait_host(k);
    e = cudaMemcpyAsync(S[issued].d, R.at(k, S[issued].h), S[issued].n, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) e = cudaEventRecord(R.ev[k], st);
    R.dma_pending[k] = true;
    ++issued;
  }
  for (size_t p = issued; p < dispatched; ++p) R.wait_host((int)((base + p) % kSlots));
  MVS_REQUIRE(e == cudaSuccess, MVS_ERR_CUDA, "staged upload failed: %s", cudaGetErrorString(e));
  return MVS_OK;
}

static int staged_download(char* h_dst, size_t h_pitch, size_t h_plane, const char* d_src, size_t d_pitch,
                           size_t d_plane, size_t width, size_t rows, size_t planes, cudaStream_t st) {
  if (width == 0 || rows == 0 || planes == 0) return MVS_OK;
  MVS_REQUIRE(h_dst && d_src, MVS_ERR_INVALID, "NULL pointer");
  MVS_REQUIRE(d_pitch >= width && h_pitch >= width, MVS_ERR_INVALID, "pitch smaller than width");
  int dev = 0;
  MVS_CHECK_CUDA(cudaGetDevice(&dev));
  const Pieces P(width, rows, planes, d_pitch, d_plane, h_pitch, h_plane);
  Ring& R = ring(1);
  std::lock_guard<std::mutex> lock(R.mtx);
  MVS_CHECK_CUDA(R.prepare(dev, P.buf_bytes));
  // DMAs run up to kDmaAhead pieces ahead on `stream`; a piece that has landed is drained to the
  // user's array by a pool thread (cache-bypassing stores) and its slot is reused afterwards
  size_t issued = 0, drained = 0;  // drained = pieces handed to the pool
  bool draining[kSlots] = {};
  cudaError_t e = cudaSuccess;
  while (drained < P.n && e == cudaSuccess) {
    while (issued < P.n && issued < drained + kDmaAhead) {
      const int k = (int)(issued % kSlots);
      if (draining[k]) {
        R.wait_host(k);
        draining[k] = false;
      }
      size_t d_off, h_off, n_rows, bytes;
      P.locate(issued, d_pitch, d_plane, h_pitch, h_plane, &d_off, &h_off, &n_rows, &bytes);
      char* stage = R.at(k, h_dst + h_off);
      if (n_rows == 0) e = cudaMemcpyAsync(stage, d_src + d_off, bytes, cudaMemcpyDeviceToHost, st);
      else e = cudaMemcpy2DAsync(stage, width, d_src + d_off, d_pitch, width, n_rows, cudaMemcpyDeviceToHost, st);
      if (e == cudaSuccess) e = cudaEventRecord(R.ev[k], st);
      if (e != cudaSuccess) break;
      ++issued;
    }
    if (e != cudaSuccess) break;
    const int k = (int)(drained % kSlots);
    if ((e = cudaEventSynchronize(R.ev[k])) != cudaSuccess) break;
    R.clear(k);
    draining[k] = true;
    const size_t p = drained++;
    pool().submit([&R, &P, k, p, h_dst, d_pitch, d_plane, h_pitch, h_plane, width] {
      size_t d_off, h_off, n_rows, bytes;
      P.locate(p, d_pitch, d_plane, h_pitch, h_plane, &d_off, &h_off, &n_rows, &bytes);
      const char* stage = R.at(k, h_dst + h_off);
      if (n_rows == 0) copy_stream_stores(h_dst + h_off, stage, bytes);
      else rows_copy<true>(h_dst + h_off, h_pitch, stage, width, width, n_rows);
#if defined(__SSE2__)
      _mm_sfence();
#endif
      R.mark(k);
    });
  }
  for (int k = 0; k < kSlots; ++k)
    if (draining[k]) R.wait_host(k);
  if (e != cudaSuccess) cudaStreamSynchronize(st);  // pending DMAs must not write the ring after an error return
  for (int k = 0; k < kSlots; ++k) R.dma_pending[k] = false;
  MVS_REQUIRE(e == cudaSuccess, MVS_ERR_CUDA, "staged download failed: %s", cudaGetErrorString(e));
  return MVS_OK;
}

}  // namespace mvs

using namespace mvs;

extern "C" int mvs_copy_h2d_2d(void* d_dst, size_t d_pitch, const void* h_src, size_t h_pitch,
                               size_t width, size_t rows, void* stream) {
  return staged_upload((char*)d_dst, d_pitch, 0, (const char*)h_src, h_pitch, 0, width, rows, 1, (cudaStream_t)stream);
}

extern "C" int mvs_copy_d2h_2d(void* h_dst, size_t h_pitch, const void* d_src, size_t d_pitch,
                               size_t width, size_t rows, void* stream) {
  return staged_download((char*)h_dst, h_pitch, 0, (const char*)d_src, d_pitch, 0, width, rows, 1, (cudaStream_t)stream);
}

extern "C" int mvs_copy_h2d_many(int n, void* const* d_dst, const void* const* h_src, const size_t* bytes,
                                 void* stream) {
  if (n <= 0) return MVS_OK;
  MVS_REQUIRE(d_dst && h_src && bytes, MVS_ERR_INVALID, "NULL pointer");
  return staged_upload_many(n, (char* const*)d_dst, (const char* const*)h_src, bytes, (cudaStream_t)stream);
}

extern "C" int mvs_copy_h2d_3d(void* d_dst, size_t d_pitch, size_t d_plane, const void* h_src, size_t h_pitch,
                               size_t h_plane, size_t width, size_t rows, size_t planes, void* stream) {
  return staged_upload((char*)d_dst, d_pitch, d_plane, (const char*)h_src, h_pitch, h_plane, width, rows, planes,
                       (cudaStream_t)stream);
}

extern "C" int mvs_copy_d2h_3d(void* h_dst, size_t h_pitch, size_t h_plane, const void* d_src, size_t d_pitch,
                               size_t d_plane, size_t width, size_t rows, size_t planes, void* stream) {
  return staged_download((char*)h_dst, h_pitch, h_plane, (const char*)d_src, d_pitch, d_plane, width, rows, planes,
                         (cudaStream_t)stream);
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

// VEC = bytes per thread along x on the packed side (16, or the item size); DV: the dense side is
// 16-byte aligned too (one vector access), else it is touched item by item (ES bytes each) while the
// packed side still moves as one 16-byte vector
template <int VEC, bool PACK, bool DV = true, int ES = 1>
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
    if (VEC == 16 && !DV) {
      union { uint4 v; unsigned char b[16]; } u16;
      if (PACK) {
        u16.v = make_uint4(0, 0, 0, 0);
        if (in_zy) {
#pragma unroll
          for (int i = 0; i < 16 / ES; ++i) {
            if (xb + (i + 1) * ES <= row_bytes) {
              if (ES == 1) u16.b[i] = *reinterpret_cast<const unsigned char*>(d + i);
              else if (ES == 2) reinterpret_cast<unsigned short*>(u16.b)[i] = *reinterpret_cast<const unsigned short*>(d + 2 * i);
              else if (ES == 4) reinterpret_cast<unsigned*>(u16.b)[i] = *reinterpret_cast<const unsigned*>(d + 4 * i);
              else reinterpret_cast<unsigned long long*>(u16.b)[i] = *reinterpret_cast<const unsigned long long*>(d + 8 * i);
            }
          }
        }
        *reinterpret_cast<uint4*>(p) = u16.v;
      } else if (in_zy && xb < row_bytes) {
        u16.v = *reinterpret_cast<const uint4*>(p);
#pragma unroll
        for (int i = 0; i < 16 / ES; ++i) {
          if (xb + (i + 1) * ES <= row_bytes) {
            if (ES == 1) *reinterpret_cast<unsigned char*>(d + i) = u16.b[i];
            else if (ES == 2) *reinterpret_cast<unsigned short*>(d + 2 * i) = reinterpret_cast<unsigned short*>(u16.b)[i];
            else if (ES == 4) *reinterpret_cast<unsigned*>(d + 4 * i) = reinterpret_cast<unsigned*>(u16.b)[i];
            else *reinterpret_cast<unsigned long long*>(d + 8 * i) = reinterpret_cast<unsigned long long*>(u16.b)[i];
          }
        }
      }
    } else if (VEC == 16) {
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
  const bool pvec = ((int64_t)chunk[2] * item_size) % 16 == 0 && ((uintptr_t)packed % 16) == 0;
  const bool dvec = g.stride[0] % 16 == 0 && g.stride[1] % 16 == 0 && ((uintptr_t)dense % 16) == 0;
  const bool ialigned = ((uintptr_t)dense % item_size) == 0;
  const int64_t units = n_bytes / (pvec && (dvec || ialigned) ? 16 : item_size);
  const int blocks = (int)std::min<int64_t>((units + 255) / 256, 148 * 32);
  if (pvec && dvec) {
    chunks_copy_kernel<16, PACK><<<blocks, 256, 0, st>>>((char*)dense, (char*)packed, g, units);
  } else if (pvec && ialigned) {
    // rows of arbitrary length (e.g. 9015 floats): 16-byte vectors on the packed side only
    if (item_size == 1) chunks_copy_kernel<16, PACK, false, 1><<<blocks, 256, 0, st>>>((char*)dense, (char*)packed, g, units);
    else if (item_size == 2) chunks_copy_kernel<16, PACK, false, 2><<<blocks, 256, 0, st>>>((char*)dense, (char*)packed, g, units);
    else if (item_size == 4) chunks_copy_kernel<16, PACK, false, 4><<<blocks, 256, 0, st>>>((char*)dense, (char*)packed, g, units);
    else chunks_copy_kernel<16, PACK, false, 8><<<blocks, 256, 0, st>>>((char*)dense, (char*)packed, g, units);
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

// Pinned slots of one chunk each for the chunk store / load pipeline.  As with the staged copies
// above, only the calling thread talks to the driver; the pool's threads do the file I/O.
constexpr int kChunkSlots = 8;

struct ChunkRing {
  std::mutex mtx;
  void* buf[kChunkSlots] = {};
  size_t cap = 0;
  cudaEvent_t ev[kChunkSlots] = {};
  int device = -1;
  std::atomic<int> host_done[kChunkSlots] = {};
  cudaError_t prepare(int dev, size_t bytes) {
    cudaError_t e;
    if (device != dev) {
      for (int i = 0; i < kChunkSlots; ++i) {
        if (ev[i]) cudaEventDestroy(ev[i]);
        if ((e = cudaEventCreateWithFlags(&ev[i], cudaEventDisableTiming)) != cudaSuccess) return e;
      }
      device = dev;
    }
    if (cap < bytes) {
      for (int i = 0; i < kChunkSlots; ++i) {
        if (buf[i]) cudaFreeHost(buf[i]);
        buf[i] = nullptr;
      }
      cap = 0;
      for (int i = 0; i < kChunkSlots; ++i)
        if ((e = cudaHostAlloc(&buf[i], bytes, cudaHostAllocPortable)) != cudaSuccess) return e;
      cap = bytes;
    }
    return cudaSuccess;
  }
  void wait_host(int k) {
    for (int spins = 0; host_done[k].load(std::memory_order_acquire) == 0; ++spins) {
      if (spins < (1 << 16)) cpu_relax();
      else std::this_thread::yield();
    }
  }
};

static ChunkRing& chunk_ring() {
  static ChunkRing r;
  return r;
}

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
  cudaStream_t st = (cudaStream_t)stream;
  ChunkRing& R = chunk_ring();
  std::lock_guard<std::mutex> lock(R.mtx);
  MVS_CHECK_CUDA(R.prepare(dev, chunk_bytes));
  constexpr int kAhead = 3;  // DMAs in flight; the other slots are being written out by the pool
  std::atomic<int> failed(-1);
  bool writing[kChunkSlots] = {};
  int issued = 0, handed = 0;
  cudaError_t e = cudaSuccess;
  while (handed < n && e == cudaSuccess) {
    while (issued < n && issued < handed + kAhead) {
      const int k = issued % kChunkSlots;
      if (writing[k]) {
        R.wait_host(k);
        writing[k] = false;
      }
      e = cudaMemcpyAsync(R.buf[k], (const char*)d_packed + (size_t)issued * chunk_bytes, chunk_bytes,
                          cudaMemcpyDeviceToHost, st);
      if (e == cudaSuccess) e = cudaEventRecord(R.ev[k], st);
      if (e != cudaSuccess) break;
      ++issued;
    }
    if (e != cudaSuccess) break;
    const int k = handed % kChunkSlots;
    if ((e = cudaEventSynchronize(R.ev[k])) != cudaSuccess) break;
    R.host_done[k].store(0, std::memory_order_relaxed);
    writing[k] = true;
    const int i = handed++;
    pool().submit([&R, &failed, k, i, paths, chunk_bytes] {
      if (!file_rw(paths[i], (char*)R.buf[k], chunk_bytes, true)) {
        int expect = -1;
        failed.compare_exchange_strong(expect, i);
      }
      R.host_done[k].store(1, std::memory_order_release);
    });
  }
  for (int k = 0; k < kChunkSlots; ++k)
    if (writing[k]) R.wait_host(k);
  if (e != cudaSuccess) cudaStreamSynchronize(st);
  MVS_REQUIRE(e == cudaSuccess, MVS_ERR_CUDA, "chunk download failed: %s", cudaGetErrorString(e));
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
  cudaStream_t st = (cudaStream_t)stream;
  ChunkRing& R = chunk_ring();
  std::lock_guard<std::mutex> lock(R.mtx);
  MVS_CHECK_CUDA(R.prepare(dev, chunk_bytes));
  constexpr int kReadAhead = 6;  // files being read by the pool; the other slots hold DMAs in flight
  std::atomic<int> failed(-1);
  int status[kChunkSlots] = {};  // written by the reader of the slot: 1 data, 2 missing file
  bool dma_pending[kChunkSlots] = {};
  int dispatched = 0, issued = 0;
  cudaError_t e = cudaSuccess;
  while (issued < n && e == cudaSuccess) {
    while (dispatched < n && dispatched < issued + kReadAhead) {
      const int k = dispatched % kChunkSlots;
      if (dma_pending[k]) {
        if ((e = cudaEventSynchronize(R.ev[k])) != cudaSuccess) break;
        dma_pending[k] = false;
      }
      R.host_done[k].store(0, std::memory_order_relaxed);
      const int i = dispatched++;
      pool().submit([&R, &failed, &status, k, i, paths, chunk_bytes] {
        if (access(paths[i], F_OK) != 0) {
          status[k] = 2;  // a missing chunk reads as the fill value
        } else if (file_rw(paths[i], (char*)R.buf[k], chunk_bytes, false)) {
          status[k] = 1;
        } else {
          status[k] = 2;
          int expect = -1;
          failed.compare_exchange_strong(expect, i);
        }
        R.host_done[k].store(1, std::memory_order_release);
      });
    }
    if (e != cudaSuccess) break;
    const int k = issued % kChunkSlots;
    R.wait_host(k);
    char* dst = (char*)d_packed + (size_t)issued * chunk_bytes;
    if (status[k] == 1) {
      e = cudaMemcpyAsync(dst, R.buf[k], chunk_bytes, cudaMemcpyHostToDevice, st);
      if (e == cudaSuccess) e = cudaEventRecord(R.ev[k], st);
      dma_pending[k] = true;
    } else {
      e = cudaMemsetAsync(dst, 0, chunk_bytes, st);
    }
    ++issued;
  }
  for (int p = issued; p < dispatched; ++p) R.wait_host(p % kChunkSlots);
  // the ring is reused by the next call: its DMAs must have read the slots by then
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  MVS_REQUIRE(e == cudaSuccess, MVS_ERR_CUDA, "chunk upload failed: %s", cudaGetErrorString(e));
  const int f = failed.load();
  MVS_REQUIRE(f < 0, MVS_ERR_INVALID, "read failed for %s (short or unreadable chunk file)", paths[f]);
  return MVS_OK;
}
