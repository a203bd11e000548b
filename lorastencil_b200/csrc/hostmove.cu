// hostmove.cu -- see hostmove.h.  Host-side C++ only (no kernels).
#include "hostmove.h"

#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

namespace lora {

namespace {

constexpr size_t kPiece = 8u << 20;  // bytes per staging slot: DMA pieces of 8 MB keep both PCIe directions streaming
constexpr int kH2dSlots = 6;
constexpr int kD2hSlotsMax = 64;     // grown on demand: up to 512 MB of pinned staging for results on their way out

// a tiny fork-join pool for parallel memcpy; callable from several threads at once
class CopyPool {
  public:
    CopyPool() {
        unsigned hw = std::thread::hardware_concurrency();
        int n = (int)std::min(8u, std::max(2u, hw / 2));
        if (const char *e = getenv("LORA_COPY_THREADS")) {
            const int v = atoi(e);
            if (v >= 1 && v <= 64) n = v;
        }
        nthreads_ = n;
        for (int i = 0; i < n - 1; i++) workers_.emplace_back([this] { loop(); });
    }
    ~CopyPool() {
        {
            std::lock_guard<std::mutex> lk(m_);
            stop_ = true;
        }
        cv_.notify_all();
        for (auto &t : workers_) t.join();
    }
    void copy(void *dst, const void *src, size_t bytes) {
        if (bytes < (1u << 20) || nthreads_ == 1) {
            std::memcpy(dst, src, bytes);
            return;
        }
        const size_t part = ((bytes + nthreads_ - 1) / nthreads_ + 4095) & ~size_t(4095);
        std::atomic<int> left{0};
        std::vector<Task> mine;
        for (size_t off = 0; off < bytes; off += part)
            mine.push_back(Task{static_cast<char *>(dst) + off, static_cast<const char *>(src) + off, std::min(part, bytes - off), &left});
        left.store((int)mine.size());
        {
            std::lock_guard<std::mutex> lk(m_);
            for (size_t i = 1; i < mine.size(); i++) q_.push_back(mine[i]);
        }
        cv_.notify_all();
        run(mine[0]);
        // help with whatever is queued (ours or another caller's), then wait for our stragglers
        for (;;) {
            Task t;
            {
                std::lock_guard<std::mutex> lk(m_);
                if (q_.empty()) break;
                t = q_.front();
                q_.pop_front();
            }
            run(t);
        }
        while (left.load(std::memory_order_acquire) > 0) std::this_thread::yield();
    }

  private:
    struct Task {
        char *dst;
        const char *src;
        size_t n;
        std::atomic<int> *left;
    };
    static void run(const Task &t) {
        std::memcpy(t.dst, t.src, t.n);
        t.left->fetch_sub(1, std::memory_order_release);
    }
    void loop() {
        for (;;) {
            Task t;
            {
                std::unique_lock<std::mutex> lk(m_);
                cv_.wait(lk, [this] { return stop_ || !q_.empty(); });
                if (stop_ && q_.empty()) return;
                t = q_.front();
                q_.pop_front();
            }
            run(t);
        }
    }
    int nthreads_ = 1;
    std::vector<std::thread> workers_;
    std::mutex m_;
    std::condition_variable cv_;
    std::deque<Task> q_;
    bool stop_ = false;
};

}  // namespace

struct HostMover::Impl {
    CopyPool pool;
    // H2D: slots filled by the calling thread, freed when their DMA event has completed
    char *up[kH2dSlots] = {};
    cudaEvent_t up_ev[kH2dSlots] = {};
    int up_dev[kH2dSlots] = {};
    bool up_busy[kH2dSlots] = {};
    int up_next = 0;
    // D2H: slots filled by DMA, emptied by the drain thread
    struct Down {
        char *buf = nullptr;
        cudaEvent_t ev = nullptr;
        int dev = -1;
    };
    std::vector<Down> down;
    std::deque<int> down_free;
    struct Job {
        int slot;
        void *dst;
        size_t n;
    };
    std::deque<Job> jobs;
    size_t pending = 0;  // jobs queued or being drained
    std::mutex m;
    std::condition_variable cv_jobs, cv_free, cv_idle;
    bool stop = false;
    cudaError_t drain_err = cudaSuccess;
    std::thread drainer;

    Impl() {
        down.reserve(kD2hSlotsMax);  // the drain thread indexes `down` while the caller may append: never reallocate
        drainer = std::thread([this] { drain_loop(); });
    }
    ~Impl() {
        {
            std::lock_guard<std::mutex> lk(m);
            stop = true;
        }
        cv_jobs.notify_all();
        drainer.join();
        for (int i = 0; i < kH2dSlots; i++) {
            if (up_ev[i]) cudaEventDestroy(up_ev[i]);
            if (up[i]) cudaFreeHost(up[i]);
        }
        for (auto &d : down) {
            if (d.ev) cudaEventDestroy(d.ev);
            if (d.buf) cudaFreeHost(d.buf);
        }
    }

    void drain_loop() {
        for (;;) {
            Job j;
            {
                std::unique_lock<std::mutex> lk(m);
                cv_jobs.wait(lk, [this] { return stop || !jobs.empty(); });
                if (jobs.empty()) return;  // stop requested and nothing left
                j = jobs.front();
                jobs.pop_front();
            }
            cudaError_t e = cudaEventSynchronize(down[j.slot].ev);
            if (e == cudaSuccess) pool.copy(j.dst, down[j.slot].buf, j.n);
            {
                std::lock_guard<std::mutex> lk(m);
                if (e != cudaSuccess && drain_err == cudaSuccess) drain_err = e;
                down_free.push_back(j.slot);
                pending--;
            }
            cv_free.notify_all();
            cv_idle.notify_all();
        }
    }

    cudaError_t acquire_down(int &slot) {
        std::unique_lock<std::mutex> lk(m);
        if (down_free.empty() && (int)down.size() < kD2hSlotsMax) {
            lk.unlock();
            Down d;
            cudaError_t e = cudaHostAlloc(&d.buf, kPiece, cudaHostAllocPortable);
            lk.lock();
            if (e == cudaSuccess) {
                down.push_back(d);
                down_free.push_back((int)down.size() - 1);
            } else if (down.empty()) {
                return e;
            } else {
                cudaGetLastError();  // no more pinned memory: make do with the slots there are
            }
        }
        cv_free.wait(lk, [this] { return !down_free.empty(); });
        slot = down_free.front();
        down_free.pop_front();
        return cudaSuccess;
    }
};

bool HostMover::pageable(const void *p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return true;
    }
    return a.type == cudaMemoryTypeUnregistered;
}

HostMover::HostMover() : impl_(new Impl) {}
HostMover::~HostMover() { delete impl_; }

cudaError_t HostMover::h2d(void *dst_dev, const void *src_host, size_t bytes, cudaStream_t stream) {
    if (bytes == 0) return cudaSuccess;
    if (!pageable(src_host)) return cudaMemcpyAsync(dst_dev, src_host, bytes, cudaMemcpyHostToDevice, stream);
    Impl &I = *impl_;
    for (size_t off = 0; off < bytes; off += kPiece) {
        const size_t n = std::min(kPiece, bytes - off);
        const int s = I.up_next;
        I.up_next = (I.up_next + 1) % kH2dSlots;
        cudaError_t e;
        int dev = 0;
        cudaGetDevice(&dev);
        if (!I.up[s] && (e = cudaHostAlloc(&I.up[s], kPiece, cudaHostAllocPortable)) != cudaSuccess) return e;
        if (I.up_busy[s] && (e = cudaEventSynchronize(I.up_ev[s])) != cudaSuccess) return e;  // its previous DMA has read it
        I.up_busy[s] = false;
        if (!I.up_ev[s] || I.up_dev[s] != dev) {  // an event can only be recorded on a stream of its own device
            if (I.up_ev[s]) cudaEventDestroy(I.up_ev[s]);
            I.up_ev[s] = nullptr;
            if ((e = cudaEventCreateWithFlags(&I.up_ev[s], cudaEventDisableTiming)) != cudaSuccess) return e;
            I.up_dev[s] = dev;
        }
        I.pool.copy(I.up[s], static_cast<const char *>(src_host) + off, n);
        if ((e = cudaMemcpyAsync(static_cast<char *>(dst_dev) + off, I.up[s], n, cudaMemcpyHostToDevice, stream)) != cudaSuccess) return e;
        if ((e = cudaEventRecord(I.up_ev[s], stream)) != cudaSuccess) return e;
        I.up_busy[s] = true;
    }
    return cudaSuccess;
}

cudaError_t HostMover::d2h(void *dst_host, const void *src_dev, size_t bytes, cudaStream_t stream) {
    if (bytes == 0) return cudaSuccess;
    if (!pageable(dst_host)) return cudaMemcpyAsync(dst_host, src_dev, bytes, cudaMemcpyDeviceToHost, stream);
    Impl &I = *impl_;
    for (size_t off = 0; off < bytes; off += kPiece) {
        const size_t n = std::min(kPiece, bytes - off);
        int s = -1;
        cudaError_t e = I.acquire_down(s);
        if (e != cudaSuccess) return e;
        int dev = 0;
        cudaGetDevice(&dev);
        if (I.down[s].dev != dev) {  // the slot is free, so its event is idle: re-create it on this device
            if (I.down[s].ev) cudaEventDestroy(I.down[s].ev);
            I.down[s].ev = nullptr;
            e = cudaEventCreateWithFlags(&I.down[s].ev, cudaEventDisableTiming);
            I.down[s].dev = dev;
        }
        if (e == cudaSuccess)
            e = cudaMemcpyAsync(I.down[s].buf, static_cast<const char *>(src_dev) + off, n, cudaMemcpyDeviceToHost, stream);
        if (e == cudaSuccess) e = cudaEventRecord(I.down[s].ev, stream);
        {
            std::lock_guard<std::mutex> lk(I.m);
            if (e != cudaSuccess) {
                I.down_free.push_back(s);
            } else {
                I.jobs.push_back(Impl::Job{s, static_cast<char *>(dst_host) + off, n});
                I.pending++;
            }
        }
        if (e != cudaSuccess) return e;
        I.cv_jobs.notify_one();
    }
    return cudaSuccess;
}

HostMover &global_mover() {
    static HostMover *m = new HostMover;
    return *m;
}

cudaError_t HostMover::finish() {
    Impl &I = *impl_;
    std::unique_lock<std::mutex> lk(I.m);
    I.cv_idle.wait(lk, [&I] { return I.pending == 0; });
    const cudaError_t e = I.drain_err;
    I.drain_err = cudaSuccess;
    return e;
}

}  // namespace lora
