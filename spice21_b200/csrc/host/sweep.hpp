// Multi-GPU sweep: ONE process, one host thread + one CUDA stream per GPU, a fixed batch of B independent circuit
// instances split into contiguous blocks of ceil(B / G) per device. Instances never exchange data during a
// solve, so the only multi-GPU step is the gather of results at the end: every device copies its block of x rows /
// waveforms / status / iteration counts straight into its slice of ONE pinned, portable host buffer (all copies run
// concurrently, one DMA engine per GPU), which is what the caller reads. No collective is needed for that and none is
// invented; the reference has no multi-device path at all (it is single-threaded, analysis.rs:383-388).
//
// Each shard is an ordinary Batch (host/batch.hpp) bound to its device; a persistent worker thread per shard runs the
// shard's calls (CUDA's current device is per thread), the calling thread posts a job to every worker and waits.
#pragma once
#include <condition_variable>
#include <functional>
#include <memory>
#include <mutex>
#include <thread>

#include "batch.hpp"

namespace s21 {

// Contiguous block partition: shard g of G owns instances [first, first + count) of B; blocks of ceil(B / G), the last
// ones short or empty.
inline void sweep_partition(size_t B, int G, int g, size_t* first, size_t* count) {
  const size_t per = G > 0 ? (B + (size_t)G - 1) / (size_t)G : B;
  const size_t lo = std::min(B, per * (size_t)g), hi = std::min(B, lo + per);
  *first = lo;
  *count = hi - lo;
}

class Sweep {
 public:
  Sweep(const CktSpec& spec, const FlatCkt& flat, const std::vector<int>& devices, size_t B) : B_(B), N_(flat.n_vars()) {
    if (devices.empty()) throw S21Error(ST_OTHER, "a sweep needs at least one device");
    if (B == 0) throw S21Error(ST_OTHER, "batch size must be positive");
    const int G = (int)devices.size();
    shards_.resize((size_t)G);
    for (int g = 0; g < G; g++) {
      Shard& s = shards_[(size_t)g];
      s.device = devices[(size_t)g];
      sweep_partition(B, G, g, &s.first, &s.count);
      s.worker.reset(new Worker());
    }
    // every shard keeps a Batch (an empty block still serves the frequency-sharded AC sweep with one instance)
    run_all([&](Shard& s) { s.batch.reset(new Batch(spec, flat, s.device, std::max<size_t>(s.count, 1))); });
  }
  ~Sweep() {
    for (Shard& s : shards_) {
      if (!s.worker) continue;
      s.worker->post([&s] { s.batch.reset(); });
      s.worker->wait();
    }
    if (pinned_) cudaFreeHost(pinned_);
  }
  size_t B() const { return B_; }
  int N() const { return N_; }
  int n_devices() const { return (int)shards_.size(); }
  void shard_range(int g, int* device, size_t* first, size_t* count) const {
    const Shard& s = shards_.at((size_t)g);
    *device = s.device; *first = s.first; *count = s.count;
  }

  void add_override(const std::string& spec, const double* values) {
    run_all([&](Shard& s) {
      if (s.count) { s.batch->add_override(spec, values + s.first); return; }
      std::vector<double> one(1, values[0]);  // idle shard: keeps the circuit's parameters consistent for AC
      s.batch->add_override(spec, one.data());
    });
  }
  size_t sync_params(bool force) {
    size_t bytes = 0;
    std::mutex mu;
    run_all([&](Shard& s) {
      if (!s.count) return;
      s.batch->sync_params(force);
      std::lock_guard<std::mutex> l(mu);
      bytes += s.batch->last_h2d_bytes();
    });
    return bytes;
  }
  void reset() { for (Shard& s : shards_) s.batch->reset(); }  // lazy flag only: no CUDA call

  // dcop of all B instances; results gathered into the sweep's pinned buffer, pointers returned (valid until the next solve)
  void dcop_view(const double** x, const int32_t** status, const int32_t** iters) {
    ensure_pinned((size_t)N_ * B_ * sizeof(double) + 3 * B_ * sizeof(int32_t) * 2);
    double* X = reinterpret_cast<double*>(pinned_);
    int32_t* tails = reinterpret_cast<int32_t*>(X + (size_t)N_ * B_);       // per shard [status|iters|loads], 3 * count each
    int32_t* ST = tails + 3 * B_;
    int32_t* IT = ST + B_;
    run_all([&](Shard& s) {
      if (!s.count) return;
      s.batch->set_result_target(X + s.first * (size_t)N_, tails + 3 * s.first);
      const double* hx; const int32_t *hs, *hi;
      s.batch->dcop_device();
      s.batch->read_view(true, &hx, &hs, &hi);
      std::memcpy(ST + s.first, hs, s.count * sizeof(int32_t));
      std::memcpy(IT + s.first, hi, s.count * sizeof(int32_t));
      s.batch->set_result_target(nullptr, nullptr);
    });
    if (x) *x = X;
    if (status) *status = ST;
    if (iters) *iters = IT;
  }
  void dcop(double* x, int32_t* status, int32_t* iters) {
    const double* hx; const int32_t *hs, *hi;
    dcop_view(&hx, &hs, &hi);
    if (x) std::memcpy(x, hx, (size_t)N_ * B_ * sizeof(double));
    if (status) std::memcpy(status, hs, B_ * sizeof(int32_t));
    if (iters) std::memcpy(iters, hi, B_ * sizeof(int32_t));
  }
  void tran(double tstep, int T, const int32_t* save_vars, size_t n_save, double* wave, int32_t* status, int64_t* iters) {
    run_all([&](Shard& s) {
      if (!s.count) return;
      s.batch->tran(tstep, T, save_vars, n_save, wave ? wave + s.first * (size_t)T * n_save : nullptr, status ? status + s.first : nullptr,
                    iters ? iters + s.first : nullptr);
    });
  }
  // AC: the frequency axis is the batch; contiguous blocks of ceil(F / G) points per device
  void ac(const double* freqs, size_t F, double* x, int32_t* status, int32_t* iters) {
    const int G = (int)shards_.size();
    run_all([&](Shard& s) {
      size_t f0, fc;
      sweep_partition(F, G, (int)(&s - shards_.data()), &f0, &fc);
      if (!fc) return;
      s.batch->ac(freqs + f0, fc, x ? x + f0 * (size_t)N_ * 2 : nullptr, status ? status + f0 : nullptr, iters ? iters + f0 : nullptr);
    });
  }
  // [0] launches (sum), [1] device ms (max over devices), [2] iterations (sum), [3] loads (sum), [4..7] as Batch::stats of shard 0
  void stats(double* out8) const {
    double acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    bool first = true;
    for (const Shard& s : shards_) {
      if (!s.batch) continue;
      double o[8];
      s.batch->stats(o);
      if (first) { for (int k = 4; k < 8; k++) acc[k] = o[k]; first = false; }
      if (!s.count && o[2] == 0) continue;
      acc[0] += o[0]; acc[1] = std::max(acc[1], o[1]); acc[2] += o[2]; acc[3] += o[3];
    }
    std::memcpy(out8, acc, sizeof acc);
  }
  const char* kernel_name() const { return shards_[0].batch ? shards_[0].batch->kernel_name() : ""; }

 private:
  struct Worker {  // one persistent host thread; jobs run in order, wait() rethrows the job's exception text
    std::thread th;
    std::mutex mu;
    std::condition_variable cv;
    std::function<void()> job;
    bool busy = false, quit = false;
    std::string error;
    int error_code = 0;
    Worker() {
      th = std::thread([this] {
        std::unique_lock<std::mutex> l(mu);
        for (;;) {
          cv.wait(l, [this] { return quit || (busy && job); });
          if (quit) return;
          std::function<void()> j = std::move(job);
          job = nullptr;
          l.unlock();
          std::string err;
          int code = 0;
          try { j(); } catch (const S21Error& e) { err = e.what(); code = e.code; } catch (const std::exception& e) { err = e.what(); code = ST_OTHER; }
          l.lock();
          error = err; error_code = code; busy = false;
          cv.notify_all();
        }
      });
    }
    ~Worker() {
      { std::lock_guard<std::mutex> l(mu); quit = true; }
      cv.notify_all();
      if (th.joinable()) th.join();
    }
    void post(std::function<void()> j) {
      std::lock_guard<std::mutex> l(mu);
      job = std::move(j); busy = true; error.clear(); error_code = 0;
      cv.notify_all();
    }
    void wait() {
      std::unique_lock<std::mutex> l(mu);
      cv.wait(l, [this] { return !busy; });
    }
  };
  struct Shard {
    int device = 0;
    size_t first = 0, count = 0;
    std::unique_ptr<Batch> batch;
    std::unique_ptr<Worker> worker;
  };
  template <class F> void run_all(F f) {
    for (Shard& s : shards_) s.worker->post([&f, &s] { f(s); });
    std::string err;
    int code = 0;
    for (Shard& s : shards_) {
      s.worker->wait();
      if (err.empty() && !s.worker->error.empty()) { err = "device " + std::to_string(s.device) + ": " + s.worker->error; code = s.worker->error_code; }
    }
    if (!err.empty()) throw S21Error(code, err);
  }
  void ensure_pinned(size_t bytes) {
    if (bytes <= pinned_bytes_) return;
    if (pinned_) cudaFreeHost(pinned_);
    pinned_ = nullptr; pinned_bytes_ = 0;
    S21_CUDA(cudaHostAlloc(&pinned_, bytes, cudaHostAllocPortable));  // portable: every device's DMA may target it
    pinned_bytes_ = bytes;
  }
  size_t B_;
  int N_;
  std::vector<Shard> shards_;
  void* pinned_ = nullptr;
  size_t pinned_bytes_ = 0;
};

}  // namespace s21
