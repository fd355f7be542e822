// fdtd_staged.cu -- upload, time loop and download of a Kernel_* call as ONE pipeline (fdtd_b200_plan_run_staged):
// chunked copies, a time loop skewed along x, finished planes on their way back while later chunks still arrive;
// pageable caller arrays go through a ring of pinned bounce buffers filled by host threads.
//
// Replaces the three back-to-back phases of the reference's wrappers (cuda.cu:204-214,232-270,317-320;
// cuda_optimized.cu:308-330,402-460,476-500).
#include "fdtd_plan.h"

#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <deque>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

using namespace fdtd;

// ---------------------------------------------------------------------------- pinned bounce buffers
// Callers like the reference driver hand over pageable arrays (new float[], main.cpp:345-346): the driver would stage
// every cudaMemcpyAsync itself, synchronously, and the staged run below would fall back to one phase after the
// other.  Instead host threads copy chunk by chunk between the caller's arrays and a small ring of pinned buffers,
// and the DMA engines work on those.  The buffers are cached per process (page-locking 0.5 GB costs more than a run).
namespace {
struct StagingCache {
    std::mutex mu;
    void *up[3] = {nullptr, nullptr, nullptr};
    void *dn[2] = {nullptr, nullptr};
    size_t up_bytes = 0, dn_bytes = 0;
    bool busy = false;
} g_staging;

int staging_acquire(size_t up_bytes, size_t dn_bytes)
{
    std::lock_guard<std::mutex> lk(g_staging.mu);
    if (g_staging.busy) return (int)cudaErrorNotReady;
    if (g_staging.up_bytes < up_bytes) {
        for (void *&q : g_staging.up) {
            cudaFreeHost(q);
            q = nullptr;
        }
        g_staging.up_bytes = 0;
        for (void *&q : g_staging.up) FDTD_CHECK(cudaHostAlloc(&q, up_bytes, cudaHostAllocDefault));
        g_staging.up_bytes = up_bytes;
    }
    if (g_staging.dn_bytes < dn_bytes) {
        for (void *&q : g_staging.dn) {
            cudaFreeHost(q);
            q = nullptr;
        }
        g_staging.dn_bytes = 0;
        for (void *&q : g_staging.dn) FDTD_CHECK(cudaHostAlloc(&q, dn_bytes, cudaHostAllocDefault));
        g_staging.dn_bytes = dn_bytes;
    }
    g_staging.busy = true;
    return 0;
}
void staging_release()
{
    std::lock_guard<std::mutex> lk(g_staging.mu);
    g_staging.busy = false;
}

// memcpy split over k persistent threads (one thread moves ~10 GB/s; each PCIe direction wants ~50)
class CopyPool {
  public:
    explicit CopyPool(int k)
    {
        for (int i = 0; i < k; ++i) th_.emplace_back([this] { work(); });
    }
    ~CopyPool()
    {
        {
            std::lock_guard<std::mutex> lk(mu_);
            stop_ = true;
        }
        cv_.notify_all();
        for (auto &t : th_) t.join();
    }
    void copy(void *dst, const void *src, size_t bytes)
    {
        const int k = (int)th_.size();
        if (k <= 1 || bytes < (size_t)(1 << 20)) {
            memcpy(dst, src, bytes);
            return;
        }
        const size_t part = ((bytes + k - 1) / k + 4095) / 4096 * 4096;
        {
            std::lock_guard<std::mutex> lk(mu_);
            for (size_t off = 0; off < bytes; off += part) {
                q_.push_back(Task{(char *)dst + off, (const char *)src + off, std::min(part, bytes - off)});
                ++pending_;
            }
        }
        cv_.notify_all();
        std::unique_lock<std::mutex> lk(mu_);
        done_.wait(lk, [this] { return pending_ == 0; });
    }

  private:
    struct Task {
        char *d;
        const char *s;
        size_t n;
    };
    void work()
    {
        for (;;) {
            Task t;
            {
                std::unique_lock<std::mutex> lk(mu_);
                cv_.wait(lk, [this] { return stop_ || !q_.empty(); });
                if (q_.empty()) return;
                t = q_.front();
                q_.pop_front();
            }
            memcpy(t.d, t.s, t.n);
            {
                std::lock_guard<std::mutex> lk(mu_);
                if (--pending_ == 0) done_.notify_all();
            }
        }
    }
    std::vector<std::thread> th_;
    std::mutex mu_;
    std::condition_variable cv_, done_;
    std::deque<Task> q_;
    int pending_ = 0;
    bool stop_ = false;
};

// The halo shell of planes [xa, xb) of ONE level: every cell outside the box [X0,X1) x [Y0,Y1) x [Z0,Z1).  Planes
// outside [X0, X1) belong to it whole; of an interior plane only the y-halo rows and, per interior row, the
// (nzp - Z1) + Z0 floats that straddle the end of one row and the start of the next (contiguous in memory).
// `each(offset_floats, pitch_floats, width_floats, height)` is called for a handful of (strided) pieces, offsets
// relative to plane xa.  Pieces may overlap by a few halo floats; none touches a cell of the box.
template <class F>
void for_each_shell_piece(const fdtd::Grid &g, int xa, int xb, F &&each)
{
    const size_t plane = (size_t)g.nyp * g.nzp;
    const int lo_end = std::min(xb, std::max(xa, g.X0)), hi_begin = std::max(xa, std::min(xb, g.X1));
    if (lo_end > xa) each((size_t)0, plane, (size_t)(lo_end - xa) * plane, (size_t)1);                 // x-halo planes below the box
    if (xb > hi_begin) each((size_t)(hi_begin - xa) * plane, plane, (size_t)(xb - hi_begin) * plane, (size_t)1);  // ... above
    const int ia = lo_end, ib = hi_begin;  // interior planes of this range
    if (ib <= ia) return;
    const size_t base = (size_t)(ia - xa) * plane;
    // rows [0, Y0) of the first interior plane and the first Z0 floats of its row Y0
    each(base, plane, (size_t)g.Y0 * g.nzp + g.Z0, (size_t)1);
    // row seams: from the tail of row Y0 of plane ia to the tail of row Y1-1 of plane ib-1, every row in between
    const size_t first_row = (size_t)g.Y0, last_row = (size_t)(ib - 1 - ia) * g.nyp + (g.Y1 - 1);
    each(base + first_row * g.nzp + g.Z1, (size_t)g.nzp, (size_t)(g.nzp - g.Z1 + g.Z0), last_row - first_row + 1);
    // y-halo bands: rows [Y1, nyp) of plane x and rows [0, Y0) of plane x+1, x = ia .. ib-2
    if (ib - ia > 1) each(base + (size_t)g.Y1 * g.nzp, plane, (size_t)(g.nyp - g.Y1 + g.Y0) * g.nzp, (size_t)(ib - 1 - ia));
    // rows [Y1, nyp) of the last interior plane
    each(base + (size_t)(ib - 1 - ia) * plane + (size_t)g.Y1 * g.nzp, plane, (size_t)(g.nyp - g.Y1) * g.nzp, (size_t)1);
}

// host -> device copy of the planes [xa, xb) of one level; shell = only the halo shell (the box is write-before-read)
cudaError_t h2d_level(float *dst, const float *src, const fdtd::Grid &g, int xa, int xb, bool shell, cudaStream_t st)
{
    const size_t plane = (size_t)g.nyp * g.nzp;
    if (!shell) return cudaMemcpyAsync(dst, src, (size_t)(xb - xa) * plane * sizeof(float), cudaMemcpyHostToDevice, st);
    cudaError_t rc = cudaSuccess;
    for_each_shell_piece(g, xa, xb, [&](size_t off, size_t pitch, size_t width, size_t height) {
        if (rc != cudaSuccess) return;
        rc = height == 1 ? cudaMemcpyAsync(dst + off, src + off, width * sizeof(float), cudaMemcpyHostToDevice, st)
                         : cudaMemcpy2DAsync(dst + off, pitch * sizeof(float), src + off, pitch * sizeof(float), width * sizeof(float),
                                             height, cudaMemcpyHostToDevice, st);
    });
    return rc;
}

// the same pieces with memcpy (caller's pageable array -> pinned slot, both addressed from plane xa)
void pack_shell(float *dst, const float *src, const fdtd::Grid &g, int xa, int xb)
{
    for_each_shell_piece(g, xa, xb, [&](size_t off, size_t pitch, size_t width, size_t height) {
        for (size_t r = 0; r < height; ++r) memcpy(dst + off + r * pitch, src + off + r * pitch, width * sizeof(float));
    });
}

bool is_pinned(const void *q)
{
    cudaPointerAttributes at{};
    if (cudaPointerGetAttributes(&at, q) != cudaSuccess) {
        (void)cudaGetLastError();
        return false;
    }
    return at.type == cudaMemoryTypeHost;
}
}  // namespace

// ---------------------------------------------------------------------------- staged run (host arrays in and out)
// upload -> T time steps -> download as ONE pipeline instead of three phases (what Kernel_* does for its caller:
// cuda.cu:204-214,232-270,317-320 run them back to back, and at 512^3 the two PCIe transfers are 4x the compute).
//   * the arrays travel in chunks of B x planes (a plane is contiguous) on a copy stream;
//   * the time loop is skewed along x: block b advances ALL T steps on planes [bB - 2s, (b+1)B - 2s) for step s
//     (the stencil has radius 2, so step s of block b only needs step s-1 of blocks b and b-1): block b can run
//     as soon as chunk b+1 has landed, while later chunks are still on the wire;
//   * planes that have finished their last step go back to the host on a third stream (PCIe is full duplex),
//     x-halo planes never change and stay on the host.
// Every point sees exactly the inputs of the unskewed loop, so results are bit-identical to upload+run+download.
// section0 = device seconds of the launches of steps >= time_m+5, summed over blocks (events on the compute
// stream, waits for transfers excluded); section1 = 0 (sources are fused; halo-cell sources take the plain path).
// Returns cudaErrorNotSupported when the plain path must be taken (nothing has been touched then).
extern "C" int fdtd_b200_plan_run_staged(fdtd_b200_plan *p, float *h_u, const float *h_m, int time_m, int time_M,
                                         struct profiler *timers)
{
    if (!p || !h_u || !h_m) return (int)cudaErrorInvalidValue;
    const bool linked = p->link.peer_u[0] || p->link.peer_u[1];
    const int T = time_M - time_m + 1;
    const bool src_active = p->ncells_all > 0 && p->src_size0 > 0;
    const int nx = p->g.X1 - p->g.X0;
    int B = p->opt_stage_planes;
    if (B < 0) {
        // auto: only where the PCIe transfers dominate (a skewed loop is ~(nx + 2T)/B * T small launches), and
        // blocks long enough to keep that launch count around 1600 (the enqueue must stay ahead of the wire)
        const long long npts = (long long)nx * (p->g.Y1 - p->g.Y0) * (p->g.Z1 - p->g.Z0);
        if (npts < 8000000 || T < 1) return (int)cudaErrorNotSupported;
        B = (int)(((long long)(nx + 2 * T) * T / 1600 + 7) / 8 * 8);
        // ... and a launch should carry ~2M points: below that its fixed cost (launch, pipeline prologue) dominates and
        // the skewed loop becomes the bottleneck instead of the wire (256^3: 32 planes per block instead of 16)
        const long long per_plane = (long long)(p->g.Y1 - p->g.Y0) * (p->g.Z1 - p->g.Z0);
        B = std::max(B, (int)((2000000 / per_plane + 7) / 8 * 8));
        B = std::max(16, std::min(B, nx / 2));
    }
    if (p->shape.space_order != 4 || p->nrec_total > 0) return (int)cudaErrorNotSupported;
    if (B < 8 || linked || T < 1 || nx < 2 * B || (src_active && (p->ncells_halo > 0 || !p->opt_fuse))) return (int)cudaErrorNotSupported;
    FDTD_CHECK(cudaSetDevice(p->dev));
    if (timers) timers->section0 = timers->section1 = 0.0;
    const int saved_fuse = p->opt_t_fuse;
    p->opt_t_fuse = 1;  // one step per launch: the skew is per step (and the transfers hide the compute anyway)
    int rc = 0;
    {
        // kernel choice and tensor maps as in a plain run, but no device-side gather of mbase (m is not there yet)
        const int n2 = p->ncells2, nall = p->ncells_all;
        p->ncells2 = p->ncells_all = 0;
        rc = plan_prepare(p);
        p->ncells2 = n2;
        p->ncells_all = nall;
    }
    p->opt_t_fuse = saved_fuse;
    if (rc) return rc;
    reset_placement(p, 0);

    const Grid g = p->g;
    const size_t plane = (size_t)g.nyp * g.nzp, lvl = (size_t)g.lvl;
    // Level t2 of the first step is written over its whole box before anything reads it (every step covers
    // [X0, X1) x [Y0, Y1) x [Z0, Z1)), so only its halo shell has to travel: a quarter of the upload less.
    const int shell_level = env_int("FDTD_B200_SHELL_UPLOAD", 1) ? (((time_m + 1) % 3) + 3) % 3 : -1;
    cudaStream_t s_up = nullptr, s_down = nullptr;
    std::vector<cudaEvent_t> ev_up, ev_done, ev_t;
    std::function<void()> stop_threads = [] {};
    auto cleanup = [&](int code) {
        stop_threads();
        cudaStreamSynchronize(p->stream);
        if (s_up) cudaStreamSynchronize(s_up), cudaStreamDestroy(s_up);
        if (s_down) cudaStreamSynchronize(s_down), cudaStreamDestroy(s_down);
        for (const std::vector<cudaEvent_t> *v : {&ev_up, &ev_done, &ev_t})
            for (cudaEvent_t e : *v)
                if (e) cudaEventDestroy(e);
        return code;
    };
#define STAGED_CHECK(expr)                                   \
    do {                                                     \
        cudaError_t _e = (expr);                             \
        if (_e != cudaSuccess) return cleanup((int)_e);      \
    } while (0)
    STAGED_CHECK(cudaStreamCreateWithFlags(&s_up, cudaStreamNonBlocking));
    STAGED_CHECK(cudaStreamCreateWithFlags(&s_down, cudaStreamNonBlocking));
    // pageable caller arrays: go through the pinned bounce ring with host threads (see above)
    const int nchunks = (g.nxp + B - 1) / B;
    const bool bounce = env_int("FDTD_B200_BOUNCE", (is_pinned(h_u) && is_pinned(h_m)) ? 0 : 1) != 0;
    const int kthreads = std::max(1, std::min(env_int("FDTD_B200_COPY_THREADS", 8), (int)std::thread::hardware_concurrency() / 2));
    bool have_staging = false;
    // the bounce ring moves sub-chunks of Bb planes: slots of ~32 MB (never re-allocated when the grid changes,
    // unless one plane of the four arrays is larger than that)
    const int Bb = (int)std::max<size_t>(1, std::min<size_t>(B, ((size_t)32 << 20) / (4 * plane * sizeof(float))));
    const int nsub = (g.nxp + Bb - 1) / Bb;
    if (bounce) {
        const size_t up_bytes = std::max<size_t>((size_t)32 << 20, 4 * (size_t)Bb * plane * sizeof(float));
        const int rs = staging_acquire(up_bytes, up_bytes / 4 * 3);
        if (rs) return cleanup(rs == (int)cudaErrorNotReady ? (int)cudaErrorNotSupported : rs);
        have_staging = true;
    }
    struct DownJob {
        cudaEvent_t after;
        int x0, n;
    };
    std::mutex mu;
    std::condition_variable cv;
    int up_ready = 0;             // chunks whose H2D copies are enqueued (ev_up recorded)
    std::deque<DownJob> jobs;
    bool jobs_closed = false;
    std::atomic<int> thread_rc{0};
    std::thread t_up, t_down;
    // a worker thread failed: record the first error and wake whoever waits on cv (the main thread in
    // upload_through, the unpacking thread waiting for jobs) so the call returns the error instead of hanging
    auto fail = [&](int code) {
        {
            std::lock_guard<std::mutex> lk(mu);
            int expected = 0;
            thread_rc.compare_exchange_strong(expected, code);
        }
        cv.notify_all();
    };
    auto join_threads = [&] {
        {
            std::lock_guard<std::mutex> lk(mu);
            jobs_closed = true;
        }
        cv.notify_all();
        if (t_up.joinable()) t_up.join();
        if (t_down.joinable()) t_down.join();
        if (have_staging) staging_release(), have_staging = false;
    };
    stop_threads = [&] {
        if (t_up.joinable() || t_down.joinable()) fail((int)cudaErrorUnknown);
        join_threads();
    };
    cudaEvent_t tl[4] = {nullptr, nullptr, nullptr, nullptr};  // FDTD_B200_TRACE=1: start, H2D end, compute end, D2H end
    const bool trace = env_int("FDTD_B200_TRACE", 0) != 0;
    if (trace) {
        for (auto &e : tl) STAGED_CHECK(cudaEventCreate(&e));
        STAGED_CHECK(cudaEventRecord(tl[0], s_up));
    }

    if (src_active) {  // m at every source's base corner, straight from the host's m
        std::vector<float> mb((size_t)p->n_mbase, 1.0f);
        for (int i = 0; i < p->n_mbase; ++i)
            if (p->h_base_idx[i] >= 0) mb[i] = h_m[p->h_base_idx[i]];
        STAGED_CHECK(cudaMemcpyAsync(p->d_mbase, mb.data(), mb.size() * sizeof(float), cudaMemcpyHostToDevice, p->stream));
        STAGED_CHECK(cudaStreamSynchronize(p->stream));
    }

    // upload chunks: padded planes [c*B, (c+1)*B) of u (three levels) and m
    ev_up.resize(bounce ? nsub : nchunks);
    if (bounce) {
        for (auto &e : ev_up) STAGED_CHECK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        t_up = std::thread([&] {  // pack a sub-chunk into a pinned slot with host threads, then hand it to the DMA engine
            cudaSetDevice(p->dev);
            CopyPool pool(kthreads);
            for (int c = 0; c < nsub && !thread_rc.load(); ++c) {
                const size_t x0 = (size_t)c * Bb, n = std::min<size_t>(Bb, g.nxp - x0) * plane;
                float *slot = static_cast<float *>(g_staging.up[c % 3]);
                cudaError_t e = c >= 3 ? cudaEventSynchronize(ev_up[c - 3]) : cudaSuccess;  // the slot's previous sub-chunk has left
                const int xa = (int)x0, xb = (int)(x0 + n / plane);
                for (int r = 0; r < 4 && e == cudaSuccess; ++r) {
                    const float *srcp = r < 3 ? h_u + r * lvl + x0 * plane : h_m + x0 * plane;
                    float *dstp = r < 3 ? p->d_u + r * lvl + x0 * plane : p->d_m + x0 * plane;
                    if (r == shell_level)
                        pack_shell(slot + r * n, srcp, g, xa, xb);
                    else
                        pool.copy(slot + r * n, srcp, n * sizeof(float));
                    e = h2d_level(dstp, slot + r * n, g, xa, xb, r == shell_level, s_up);
                }
                if (e == cudaSuccess) e = cudaEventRecord(ev_up[c], s_up);
                if (e != cudaSuccess) {
                    fail((int)e);
                    break;
                }
                {
                    std::lock_guard<std::mutex> lk(mu);
                    up_ready = c + 1;
                }
                cv.notify_all();
            }
        });
        t_down = std::thread([&] {  // D2H into a pinned slot, then host threads unpack it while the next D2H runs
            cudaSetDevice(p->dev);
            CopyPool pool(kthreads);
            std::vector<cudaEvent_t> evs;
            DownJob pending{nullptr, 0, 0};
            cudaEvent_t pending_landed = nullptr;  // the pending job's own "copy has landed" event
            int j = 0, pending_slot = 0;
            auto unpack = [&](const DownJob &d, int slot_i, cudaEvent_t landed) {
                cudaError_t e = landed ? cudaEventSynchronize(landed) : cudaErrorUnknown;
                if (e != cudaSuccess) {
                    fail((int)e);
                    return;
                }
                const size_t n = (size_t)d.n * plane;
                const float *slot = static_cast<const float *>(g_staging.dn[slot_i]);
                for (int r = 0; r < 3; ++r) pool.copy(h_u + r * lvl + (size_t)d.x0 * plane, slot + r * n, n * sizeof(float));
            };
            for (;;) {
                DownJob d;
                {
                    std::unique_lock<std::mutex> lk(mu);
                    cv.wait(lk, [&] { return !jobs.empty() || jobs_closed; });
                    if (jobs.empty()) break;
                    d = jobs.front();
                    jobs.pop_front();
                }
                const int slot_i = j & 1;
                const size_t n = (size_t)d.n * plane;
                float *slot = static_cast<float *>(g_staging.dn[slot_i]);
                cudaEvent_t landed = nullptr;
                cudaError_t e = cudaStreamWaitEvent(s_down, d.after, 0);
                for (int r = 0; r < 3 && e == cudaSuccess; ++r)
                    e = cudaMemcpyAsync(slot + r * n, p->d_u + r * lvl + (size_t)d.x0 * plane, n * sizeof(float), cudaMemcpyDeviceToHost, s_down);
                if (e == cudaSuccess) e = cudaEventCreateWithFlags(&landed, cudaEventDisableTiming);
                if (e == cudaSuccess) e = cudaEventRecord(landed, s_down);
                if (e != cudaSuccess) fail((int)e);
                if (landed) evs.push_back(landed);
                if (pending.n > 0) unpack(pending, pending_slot, pending_landed);
                if (e != cudaSuccess) {
                    pending.n = 0;
                    break;
                }
                pending = d;
                pending_landed = landed;
                pending_slot = slot_i;
                ++j;
            }
            if (pending.n > 0) unpack(pending, pending_slot, pending_landed);
            for (cudaEvent_t e : evs) cudaEventDestroy(e);
        });
    }
    int uploaded = 0;
    auto upload_through = [&](int c_last) -> int {
        if (bounce) {  // the packing thread enqueues the copies; wait until sub-chunk c_last is on its way
            c_last = std::min(c_last, nsub - 1);
            std::unique_lock<std::mutex> lk(mu);
            cv.wait(lk, [&] { return up_ready > c_last || thread_rc.load() != 0; });
            return thread_rc.load();
        }
        for (; uploaded <= c_last && uploaded < nchunks; ++uploaded) {
            const size_t x0 = (size_t)uploaded * B, n = std::min<size_t>(B, g.nxp - x0) * plane;
            const int xa = (int)x0, xb = (int)(x0 + n / plane);
            for (int r = 0; r < 3; ++r)
                FDTD_CHECK(h2d_level(p->d_u + r * lvl + x0 * plane, h_u + r * lvl + x0 * plane, g, xa, xb, r == shell_level, s_up));
            FDTD_CHECK(cudaMemcpyAsync(p->d_m + x0 * plane, h_m + x0 * plane, n * sizeof(float), cudaMemcpyHostToDevice, s_up));
            FDTD_CHECK(cudaEventCreateWithFlags(&ev_up[uploaded], cudaEventDisableTiming));
            FDTD_CHECK(cudaEventRecord(ev_up[uploaded], s_up));
        }
        return 0;
    };

    // one-step launch plan for a block: the streaming kernel with one x chunk per tile
    // ... unless the block's tiles alone cannot fill the machine (256^3 with 8 x 64 tiles: 128 CTAs on 148 SMs, each
    // streaming all B planes -- latency-bound, 30 us for 5 us of work): then the block is cut into x chunks of >= 8 planes
    // until there are ~4 CTAs per SM
    TmaPlan tma_block = p->tma;
    tma_block.xchunk = B;
    if (p->kernel_used == 2) {
        const int tiles = ((g.Y1 - g.Y0 + p->tma.ty - 1) / p->tma.ty) * ((g.Z1 - g.Z0 + p->tma.tz - 1) / p->tma.tz);
        int nch = (4 * p->sm_count + tiles - 1) / tiles;
        nch = std::max(1, std::min(nch, B / 8));
        tma_block.xchunk = (B + nch - 1) / nch;
    }
    const int first_timed = time_m + FDTD_WARMUP_STEPS;
    // skewed blocks: b = 0 .. nblocks-1 until the last step's window has passed X1
    const int nblocks = (nx + 2 * (T - 1) + B - 1) / B;
    ev_done.resize(nblocks);
    int downloaded = g.X0;  // padded planes below this are final on the host
    for (int b = 0; b < nblocks && !rc; ++b) {
        // step 0 of this block reads up to padded plane X0 + (b+1)B + 1
        const int need = std::min(g.nxp - 1, g.X0 + (b + 1) * B + 1);
        const int cdiv = bounce ? Bb : B, clast = (bounce ? nsub : nchunks) - 1;
        if ((rc = upload_through(need / cdiv))) break;
        STAGED_CHECK(cudaStreamWaitEvent(p->stream, ev_up[std::min(need / cdiv, clast)], 0));
        bool timing = false;
        for (int s = 0; s < T && !rc; ++s) {
            const int time = time_m + s;
            const int lo = std::max(g.X0, g.X0 + b * B - 2 * s), hi = std::min(g.X1, g.X0 + (b + 1) * B - 2 * s);
            if (hi <= lo) continue;
            if (time >= first_timed && !timing) {
                cudaEvent_t e;
                STAGED_CHECK(cudaEventCreate(&e));
                STAGED_CHECK(cudaEventRecord(e, p->stream));
                ev_t.push_back(e);
                timing = true;
            }
            const int t0 = ((time % 3) + 3) % 3, t1 = (((time + 2) % 3) + 3) % 3, t2 = (((time + 1) % 3) + 3) % 3;
            StepArgs a{};
            a.u = p->d_u;
            a.m = p->d_m;
            a.g = g;
            a.g.X0 = lo;
            a.g.X1 = hi;
            a.k = p->k;
            a.t0 = t0;
            a.t1 = t1;
            a.t2 = t2;
            if (src_active && time >= 0 && time < p->src_size0 && p->ncells_int > 0) {
                a.sv.plane_off = p->d_plane_off;
                a.sv.cells = p->d_cells;
                a.sv.contribs = p->d_contribs;
                a.sv.src_row = p->d_src + (size_t)time * p->pstride;
                a.sv.mbase = p->d_mbase;
                a.sv.ncells = p->ncells_int;
            }
            a.link.depth = 2;
            rc = p->kernel_used == 2 ? launch_stencil_tma(tma_block, a, p->opt_exact != 0, p->stream)
                                     : launch_stencil_generic(a, p->opt_exact != 0, p->stream);
            p->last_launches++;
        }
        if (rc) break;
        if (timing) {
            cudaEvent_t e;
            STAGED_CHECK(cudaEventCreate(&e));
            STAGED_CHECK(cudaEventRecord(e, p->stream));
            ev_t.push_back(e);
        }
        STAGED_CHECK(cudaEventCreateWithFlags(&ev_done[b], cudaEventDisableTiming));
        STAGED_CHECK(cudaEventRecord(ev_done[b], p->stream));
        // planes below X0 + (b+1)B - 2(T-1) have been through all T steps
        const int fin = b == nblocks - 1 ? g.X1 : std::min(g.X1, g.X0 + (b + 1) * B - 2 * (T - 1));
        if (fin > downloaded && bounce) {
            {
                std::lock_guard<std::mutex> lk(mu);
                for (int x = downloaded; x < fin; x += Bb) jobs.push_back(DownJob{ev_done[b], x, std::min(Bb, fin - x)});
            }
            cv.notify_all();
            downloaded = fin;
        } else if (fin > downloaded) {
            STAGED_CHECK(cudaStreamWaitEvent(s_down, ev_done[b], 0));
            const size_t n = (size_t)(fin - downloaded) * plane;
            for (int r = 0; r < 3; ++r)
                STAGED_CHECK(cudaMemcpyAsync(h_u + r * lvl + (size_t)downloaded * plane, p->d_u + r * lvl + (size_t)downloaded * plane,
                                             n * sizeof(float), cudaMemcpyDeviceToHost, s_down));
            downloaded = fin;
        }
    }
    if (!rc) rc = upload_through((bounce ? nsub : nchunks) - 1);  // trailing halo planes (the device copy stays complete for later runs)
    if (bounce) {
        if (rc) fail(rc);
        join_threads();  // all chunks are on their way, all finished planes are back in the caller's array
        if (!rc) rc = thread_rc.load();
    }
    if (rc) return cleanup(rc);
    if (trace) {
        cudaEventRecord(tl[1], s_up);
        cudaEventRecord(tl[2], p->stream);
        cudaEventRecord(tl[3], s_down);
    }
    STAGED_CHECK(cudaStreamSynchronize(s_up));
    STAGED_CHECK(cudaStreamSynchronize(p->stream));
    STAGED_CHECK(cudaStreamSynchronize(s_down));
    if (trace) {
        float t1 = 0.f, t2 = 0.f, t3 = 0.f;
        cudaEventElapsedTime(&t1, tl[0], tl[1]);
        cudaEventElapsedTime(&t2, tl[0], tl[2]);
        cudaEventElapsedTime(&t3, tl[0], tl[3]);
        fprintf(stderr, "[fdtd_b200] staged: %d blocks of %d planes, %ld launches; H2D done at %.2f ms, compute at %.2f, D2H at %.2f\n",
                nblocks, B, p->last_launches, t1, t2, t3);
        for (auto &e : tl) cudaEventDestroy(e);
    }
    double s0 = 0.0;
    for (size_t i = 0; i + 1 < ev_t.size(); i += 2) {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, ev_t[i], ev_t[i + 1]);
        s0 += ms * 1e-3;
    }
    const int ntimed = time_M >= first_timed ? time_M - first_timed + 1 : 0;
    p->last_kernel_seconds = ntimed > 0 ? s0 / ntimed : 0.0;
    if (timers) timers->section0 = s0;
#undef STAGED_CHECK
    return cleanup(0);
}

