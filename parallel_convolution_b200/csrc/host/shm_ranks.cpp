// CPU ranks on one node without messages: the host-side counterpart of the GPU design.
//
// The reference's ranks own padded private tiles and trade halo rows/columns through MPI every
// iteration (mpi/mpi_conv.c:113-130, 164-248). On one node that is copying for its own sake: here all
// ranks map ONE shared segment that holds the two ping-pong images, and a rank reads the halo it needs
// straight out of its neighbours' part of the image — "peer memory" for processes, exactly what the GPU
// backends do over NVLink. What is left of the exchange is one counter per rank:
//
//   done[r] = number of epochs rank r has completed (an epoch = up to k iterations, temporal blocking).
//   Epoch e reads image[e & 1] and writes the rank's own tile into image[(e + 1) & 1].
//   A band of rank r may run epoch e once every neighbour n it touches has done[n] >= e:
//     - read-after-write: n's part of image[e & 1] is final when n has finished epoch e - 1;
//     - write-after-read: the rows r overwrites in image[(e + 1) & 1] were last read by n during epoch
//       e - 1 (same buffer, two epochs apart) — the same condition.
//   n cannot run ahead and overwrite what r still reads, because n's epoch e + 1 waits for done[r] >= e + 1.
//
// Inside a rank the tile is cut into row bands that advance k iterations in cache-sized private buffers
// (same scheme as pconv_cpu_run_image_blocked), spread over OpenMP threads (the reference's
// mpi_omp_conv, mpi_omp/mpi_omp_conv.c:297-310). Bands that do not touch a tile edge wait for nobody, so
// they overlap the neighbours' lateness the way the reference overlaps its interior pass with MPI_Irecv.
// A wait that exceeds the timeout fails the run (-ETIMEDOUT) and raises the rank's abort flag so that
// ranks waiting on it fail fast (-ECANCELED) instead of hanging: the CPU face of the in-kernel watchdog.
#include <fcntl.h>
#include <omp.h>
#include <sched.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <time.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <cerrno>
#include <cstdint>
#include <cstring>
#include <new>
#include <vector>

#include "pconv_host.h"

namespace {

constexpr uint64_t kMagic = 0x70636f6e7673686dULL;   // "pconvshm"
constexpr int64_t kPage = 4096;

struct Header {
    uint64_t magic;
    int64_t H, W, pitch, image_bytes, off_counters, off_image[2], total_bytes;
    int32_t C, nranks;
};

struct alignas(64) Counter {
    std::atomic<int64_t> done;     // epochs completed since the last reset
    std::atomic<int32_t> abort;    // set when this rank gave up: waiters stop waiting
    double value;                  // a number the rank leaves for whoever collects results (its loop time)
};

// Process barrier for launchers that have no process group (pconv_native -np N): central counter + generation.
struct alignas(64) Barrier {
    std::atomic<int64_t> count, gen;
};
constexpr int64_t kBarrierOffset = 2048;   // inside the header page, after the Header fields
static_assert(sizeof(Counter) == 64, "one cache line per rank");
static_assert(std::atomic<int64_t>::is_always_lock_free, "counters live in shared memory");

int64_t round_up(int64_t v, int64_t a) { return (v + a - 1) / a * a; }

double now_s() {
    timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

}  // namespace

struct pconv_shm {
    int fd;
    uint8_t* base;
    Header* h;
    Counter* ctr;
    Barrier* bar() const { return (Barrier*)(base + kBarrierOffset); }
};
static_assert(sizeof(Header) <= kBarrierOffset, "the barrier sits behind the header");

namespace {

pconv_shm* map_segment(int fd, int64_t bytes) {
    void* p = mmap(nullptr, (size_t)bytes, PROT_READ | PROT_WRITE, MAP_SHARED | MAP_POPULATE, fd, 0);   // prefault: no page faults inside timed runs
    if (p == MAP_FAILED) return nullptr;
    pconv_shm* s = new (std::nothrow) pconv_shm;
    if (!s) {
        munmap(p, (size_t)bytes);
        return nullptr;
    }
    s->fd = fd;
    s->base = (uint8_t*)p;
    s->h = (Header*)p;
    return s;
}

// 0 when every neighbour in `need` (bit = index into nbr[9]) has completed `epoch` epochs
int wait_neighbours(const pconv_shm* s, const int32_t* nbr, unsigned need, int64_t epoch, double timeout_s) {
    for (int d = 0; d < 9; ++d) {
        if (!(need >> d & 1u) || nbr[d] < 0) continue;
        const Counter& c = s->ctr[nbr[d]];
        double t0 = 0.0;
        for (uint32_t spin = 0; c.done.load(std::memory_order_acquire) < epoch; ++spin) {
            if (c.abort.load(std::memory_order_relaxed)) return -ECANCELED;
            if (spin < 256) {
                __builtin_ia32_pause();
            } else if (spin < 4096) {
                sched_yield();              // ranks x threads may exceed the cores
            } else {
                if (t0 == 0.0) t0 = now_s();
                if (timeout_s > 0 && now_s() - t0 > timeout_s) return -ETIMEDOUT;
                timespec nap = {0, 50000};
                nanosleep(&nap, nullptr);
            }
        }
    }
    return 0;
}

struct Band {
    int64_t r0, r1;
    unsigned need;      // neighbours whose tiles this band reads from / whose reads it overwrites
};

}  // namespace

extern "C" pconv_shm_t* pconv_shm_create(const char* path, int64_t H, int64_t W, int C, int nranks) {
    if (!path || H <= 0 || W <= 0 || C <= 0 || nranks <= 0) return nullptr;
    Header h{};
    h.magic = kMagic;
    h.H = H;
    h.W = W;
    h.C = C;
    h.nranks = nranks;
    h.pitch = round_up(W * (int64_t)C + 2 * C, 64);
    h.image_bytes = round_up((H + 2) * h.pitch, kPage);
    h.off_counters = kPage;
    h.off_image[0] = h.off_counters + round_up((int64_t)nranks * (int64_t)sizeof(Counter), kPage);
    h.off_image[1] = h.off_image[0] + h.image_bytes;
    h.total_bytes = h.off_image[1] + h.image_bytes;
    const int fd = open(path, O_CREAT | O_EXCL | O_RDWR | O_CLOEXEC, 0600);
    if (fd < 0) return nullptr;
    // posix_fallocate, not ftruncate: tmpfs must really have the room (ENOSPC now instead of SIGBUS at the first
    // touch). The pages come back zeroed: the 1-pixel frame of both images is the zero padding and is never written.
    if (posix_fallocate(fd, 0, (off_t)h.total_bytes) != 0) {
        close(fd);
        unlink(path);
        return nullptr;
    }
    pconv_shm* s = map_segment(fd, h.total_bytes);
    if (!s) {
        close(fd);
        unlink(path);
        return nullptr;
    }
    *s->h = h;
    s->ctr = (Counter*)(s->base + h.off_counters);
    for (int r = 0; r < nranks; ++r) {
        new (&s->ctr[r]) Counter;
        s->ctr[r].done.store(0, std::memory_order_relaxed);
        s->ctr[r].abort.store(0, std::memory_order_relaxed);
        s->ctr[r].value = 0.0;
    }
    new (s->bar()) Barrier;
    s->bar()->count.store(0, std::memory_order_relaxed);
    s->bar()->gen.store(0, std::memory_order_relaxed);
    std::atomic_thread_fence(std::memory_order_release);
    return s;
}

extern "C" pconv_shm_t* pconv_shm_attach(const char* path) {
    if (!path) return nullptr;
    const int fd = open(path, O_RDWR | O_CLOEXEC);
    if (fd < 0) return nullptr;
    struct stat st;
    if (fstat(fd, &st) != 0 || st.st_size < kPage) {
        close(fd);
        return nullptr;
    }
    pconv_shm* s = map_segment(fd, st.st_size);
    if (!s) {
        close(fd);
        return nullptr;
    }
    if (s->h->magic != kMagic || s->h->total_bytes != (int64_t)st.st_size) {
        munmap(s->base, (size_t)st.st_size);
        close(fd);
        delete s;
        return nullptr;
    }
    s->ctr = (Counter*)(s->base + s->h->off_counters);
    return s;
}

extern "C" void pconv_shm_close(pconv_shm_t* s) {
    if (!s) return;
    const int64_t bytes = s->h->total_bytes;
    munmap(s->base, (size_t)bytes);
    close(s->fd);
    delete s;
}

extern "C" uint8_t* pconv_shm_pixels(pconv_shm_t* s, int which) {
    if (!s) return nullptr;
    return s->base + s->h->off_image[which & 1] + s->h->pitch + s->h->C;   // pixel (0,0), inside the frame
}

extern "C" int64_t pconv_shm_pitch(const pconv_shm_t* s) { return s ? s->h->pitch : 0; }
extern "C" int64_t pconv_shm_bytes(const pconv_shm_t* s) { return s ? s->h->total_bytes : 0; }

extern "C" int64_t pconv_shm_epoch(const pconv_shm_t* s, int rank) {
    if (!s || rank < 0 || rank >= s->h->nranks) return -1;
    return s->ctr[rank].done.load(std::memory_order_acquire);
}

extern "C" void pconv_shm_reset(pconv_shm_t* s, int rank) {
    if (!s || rank < 0 || rank >= s->h->nranks) return;
    s->ctr[rank].abort.store(0, std::memory_order_relaxed);
    s->ctr[rank].done.store(0, std::memory_order_release);
}

extern "C" int pconv_shm_barrier(pconv_shm_t* s, double timeout_s) {
    if (!s) return -EINVAL;
    Barrier* b = s->bar();
    const int64_t gen = b->gen.load(std::memory_order_acquire);
    if (b->count.fetch_add(1, std::memory_order_acq_rel) + 1 == s->h->nranks) {
        b->count.store(0, std::memory_order_relaxed);
        b->gen.store(gen + 1, std::memory_order_release);
        return 0;
    }
    double t0 = 0.0;
    for (uint32_t spin = 0; b->gen.load(std::memory_order_acquire) == gen; ++spin) {
        if (spin < 1024) {
            sched_yield();
        } else {
            if (t0 == 0.0) t0 = now_s();
            if (timeout_s > 0 && now_s() - t0 > timeout_s) return -ETIMEDOUT;
            for (int r = 0; r < s->h->nranks; ++r)
                if (s->ctr[r].abort.load(std::memory_order_relaxed)) return -ECANCELED;
            timespec nap = {0, 100000};
            nanosleep(&nap, nullptr);
        }
    }
    return 0;
}

extern "C" void pconv_shm_set_value(pconv_shm_t* s, int rank, double v) {
    if (s && rank >= 0 && rank < s->h->nranks) s->ctr[rank].value = v;
}

extern "C" double pconv_shm_get_value(const pconv_shm_t* s, int rank) {
    return (s && rank >= 0 && rank < s->h->nranks) ? s->ctr[rank].value : 0.0;
}

extern "C" void pconv_shm_abort(pconv_shm_t* s, int rank) {
    if (!s || rank < 0 || rank >= s->h->nranks) return;
    s->ctr[rank].abort.store(1, std::memory_order_release);
}

extern "C" int pconv_shm_run(pconv_shm_t* s, int rank, int64_t r0, int64_t r1, int64_t c0, int64_t c1,
                             const int32_t nbr[9], const pconv_filter_t* f, int loops, int k, int threads,
                             double timeout_s, int64_t* epochs_run) {
    if (!s || !f || !nbr || rank < 0 || rank >= s->h->nranks || loops < 0 || k < 1) return -EINVAL;
    const Header& h = *s->h;
    if (r0 < 0 || r1 > h.H || r0 >= r1 || c0 < 0 || c1 > h.W || c0 >= c1) return -EINVAL;
    const int C = h.C;
    const int64_t H = h.H, W = h.W, pitch = h.pitch;
    const int nt = threads > 0 ? threads : 1;
    // The halo must come from direct neighbours only: their counters are the only ones waited on.
    for (int d = 0; d < 9; ++d)
        if (d != 4 && nbr[d] >= s->h->nranks) return -EINVAL;

    // private band geometry: every band spans the tile's columns plus a k-deep halo, clipped to the image
    const int64_t lpitch = round_up((c1 - c0 + 2 * k + 2) * C, 64);
    int64_t band = (int64_t)(1 << 20) / (2 * lpitch) - 2 * k;
    band = std::max<int64_t>(band, 8);
    band = std::min<int64_t>(band, std::max<int64_t>((r1 - r0 + nt - 1) / nt, 8));
    std::vector<Band> bands;
    const unsigned west = (c0 > 0 && nbr[3] >= 0) ? 1u << 3 : 0u, east = (c1 < W && nbr[5] >= 0) ? 1u << 5 : 0u;
    for (int64_t b0 = r0; b0 < r1; b0 += band) {
        Band b{b0, std::min(r1, b0 + band), west | east};
        if (b.r0 - k < r0 && r0 > 0) b.need |= 1u << 1 | (west ? 1u << 0 : 0u) | (east ? 1u << 2 : 0u);
        if (b.r1 + k > r1 && r1 < H) b.need |= 1u << 7 | (west ? 1u << 6 : 0u) | (east ? 1u << 8 : 0u);
        bands.push_back(b);
    }
    // bands that depend on nobody first: they hide a late neighbour
    std::stable_sort(bands.begin(), bands.end(), [](const Band& a, const Band& b) { return (a.need == 0) > (b.need == 0); });
    const int64_t nb = (int64_t)bands.size();
    const size_t lbytes = (size_t)pconv_cpu_band_scratch_bytes(band, c1 - c0, C, k);

    int64_t epoch = s->ctr[rank].done.load(std::memory_order_acquire);
    int64_t ran = 0;
    std::atomic<int> failed{0};
    std::vector<std::vector<uint8_t>> scratch((size_t)nt * 2);
    for (int done = 0; done < loops && !failed.load();) {
        const int kk = std::min(k, loops - done);
        const uint8_t* src = s->base + h.off_image[epoch & 1] + pitch + C;
        uint8_t* dst = s->base + h.off_image[(epoch + 1) & 1] + pitch + C;
#pragma omp parallel num_threads(nt) if (nt > 1)
        {
            const int tid = omp_get_thread_num();
            std::vector<uint8_t>& la = scratch[(size_t)tid * 2];
            std::vector<uint8_t>& lb = scratch[(size_t)tid * 2 + 1];
            if (la.size() < lbytes) la.assign(lbytes, 0);
            if (lb.size() < lbytes) lb.assign(lbytes, 0);
#pragma omp for schedule(dynamic, 1)
            for (int64_t i = 0; i < nb; ++i) {
                if (failed.load(std::memory_order_relaxed)) continue;
                const Band& b = bands[(size_t)i];
                if (b.need) {
                    const int rc = wait_neighbours(s, nbr, b.need, epoch, timeout_s);
                    if (rc != 0) {
                        int zero = 0;
                        failed.compare_exchange_strong(zero, rc);
                        continue;
                    }
                }
                pconv_cpu_band_advance(src, dst, pitch, H, W, C, f, b.r0, b.r1, c0, c1, kk, la.data(), lb.data());
            }
        }
        if (failed.load()) break;
        ++epoch;
        ++ran;
        s->ctr[rank].done.store(epoch, std::memory_order_release);   // publishes every band's stores
        done += kk;
    }
    if (epochs_run) *epochs_run = ran;
    const int rc = failed.load();
    if (rc != 0) s->ctr[rank].abort.store(1, std::memory_order_release);
    return rc;
}
