// CPU ranks with private tiles and messages over TCP: the distributed-memory counterpart of shm_ranks.cpp,
// for ranks that do not share a node — what `mpirun -np N ./mpi_conv` is on a cluster, without MPI.
//
// Structure of the reference's program (mpi/mpi_conv.c), re-designed rather than translated:
//   * padded ping-pong tiles (mpi_conv.c:113-130) — here with a k-deep halo, so one exchange serves k iterations
//     (temporal blocking: iteration j of an epoch recomputes a ring of k-j halo pixels; bit-identical);
//   * all EIGHT neighbours are exchanged (the reference skips the diagonals, defect D1), strips are packed into
//     contiguous messages instead of MPI_Type_vector (mpi_conv.c:80-88);
//   * the reference's schedule (mpi_conv.c:164-248: Isend/Irecv -> inner pixels -> Wait -> border pixels):
//     messages are written eagerly, the part of the epoch's first iteration that reads no halo runs while
//     they are in flight, then the halos are awaited and the frame is computed;
//   * MPI_Barrier / the max-reduction by Send/Recv (mpi_conv.c:159,273-283) go through rank 0.
// Transport: one TCP connection per pair of ranks (full mesh, TCP_NODELAY), framed messages, non-blocking
// sockets driven by poll(); whenever a write would block every socket is drained first, so head-to-head
// sends cannot deadlock. A peer that disappears or stays silent past the timeout fails the run.
#include <arpa/inet.h>
#include <fcntl.h>
#include <netdb.h>
#include <netinet/in.h>
#include <netinet/tcp.h>
#include <poll.h>
#include <sys/socket.h>
#include <time.h>
#include <unistd.h>

#include <algorithm>
#include <cerrno>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <deque>
#include <string>
#include <vector>

#include "pconv_host.h"

namespace {

constexpr int kTagBarrier = -1, kTagReduce = -2, kTagHello = -3, kTagTable = -4;

double now_s() {
    timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

struct Msg {
    int tag;
    std::vector<uint8_t> data;
};

struct Peer {
    int fd = -1;
    bool closed = false;
    std::vector<uint8_t> rx;     // bytes read, not yet a complete frame
    std::deque<Msg> q;           // complete messages nobody has taken yet
};

void set_nonblock(int fd, bool on) {
    const int fl = fcntl(fd, F_GETFL);
    fcntl(fd, F_SETFL, on ? (fl | O_NONBLOCK) : (fl & ~O_NONBLOCK));
}

void tune(int fd) {
    int one = 1;
    setsockopt(fd, IPPROTO_TCP, TCP_NODELAY, &one, sizeof one);
    int sz = 4 << 20;
    setsockopt(fd, SOL_SOCKET, SO_SNDBUF, &sz, sizeof sz);
    setsockopt(fd, SOL_SOCKET, SO_RCVBUF, &sz, sizeof sz);
}

// blocking helpers for the bootstrap only
bool write_all(int fd, const void* p, size_t n) {
    const char* c = (const char*)p;
    while (n) {
        const ssize_t w = write(fd, c, n);
        if (w > 0) { c += w; n -= (size_t)w; continue; }
        if (w < 0 && errno == EINTR) continue;
        return false;
    }
    return true;
}

bool read_all(int fd, void* p, size_t n, double deadline) {
    char* c = (char*)p;
    while (n) {
        pollfd pf{fd, POLLIN, 0};
        const double left = deadline - now_s();
        if (left <= 0) return false;
        const int pr = poll(&pf, 1, (int)std::min(left * 1e3 + 1, 1e9));
        if (pr < 0 && errno == EINTR) continue;
        if (pr <= 0) return false;
        const ssize_t r = read(fd, c, n);
        if (r > 0) { c += r; n -= (size_t)r; continue; }
        if (r < 0 && (errno == EINTR || errno == EAGAIN)) continue;
        return false;
    }
    return true;
}

}  // namespace

struct pconv_net {
    int rank = 0, world = 1;
    double timeout_s = 60.0;
    std::vector<Peer> peer;
    std::string error;

    int fail(const std::string& what) {
        error = what;
        return -EIO;
    }

    // ---- progress engine ---------------------------------------------------------------------------
    void parse(Peer& p) {
        size_t off = 0;
        while (p.rx.size() - off >= 8) {
            int32_t hdr[2];
            std::memcpy(hdr, p.rx.data() + off, 8);
            const size_t need = 8 + (size_t)hdr[1];
            if (p.rx.size() - off < need) break;
            Msg m;
            m.tag = hdr[0];
            m.data.assign(p.rx.begin() + (long)(off + 8), p.rx.begin() + (long)(off + need));
            p.q.push_back(std::move(m));
            off += need;
        }
        if (off) p.rx.erase(p.rx.begin(), p.rx.begin() + (long)off);
    }

    void drain(Peer& p) {
        uint8_t buf[1 << 16];
        for (;;) {
            const ssize_t n = read(p.fd, buf, sizeof buf);
            if (n > 0) { p.rx.insert(p.rx.end(), buf, buf + n); continue; }
            if (n == 0) { p.closed = true; break; }
            if (errno == EINTR) continue;
            if (errno != EAGAIN && errno != EWOULDBLOCK) p.closed = true;
            break;
        }
        parse(p);
    }

    // wait until something can happen (data from any peer, or room to write on out_fd); false on timeout
    bool wait_any(int out_fd, double deadline) {
        std::vector<pollfd> pf;
        std::vector<int> who;
        for (int r = 0; r < world; ++r)
            if (r != rank && peer[r].fd >= 0 && !peer[r].closed) {
                pf.push_back({peer[r].fd, POLLIN, 0});
                who.push_back(r);
            }
        if (out_fd >= 0) pf.push_back({out_fd, POLLOUT, 0});
        const double left = deadline - now_s();
        if (left <= 0) return false;
        const int n = poll(pf.data(), (nfds_t)pf.size(), (int)std::min(left * 1e3 + 1, 1e9));
        if (n < 0 && errno != EINTR) return false;
        for (size_t i = 0; i < who.size(); ++i)
            if (pf[i].revents & (POLLIN | POLLHUP | POLLERR)) drain(peer[who[i]]);
        return true;
    }

    int send(int to, int tag, const void* data, int nbytes) {
        Peer& p = peer[to];
        std::vector<uint8_t> frame(8 + (size_t)nbytes);
        const int32_t hdr[2] = {tag, nbytes};
        std::memcpy(frame.data(), hdr, 8);
        if (nbytes) std::memcpy(frame.data() + 8, data, (size_t)nbytes);
        const double deadline = now_s() + timeout_s;
        size_t off = 0;
        while (off < frame.size()) {
            const ssize_t w = write(p.fd, frame.data() + off, frame.size() - off);
            if (w > 0) { off += (size_t)w; continue; }
            if (w < 0 && errno == EINTR) continue;
            if (w < 0 && (errno == EAGAIN || errno == EWOULDBLOCK)) {
                if (!wait_any(p.fd, deadline)) return fail("timed out sending to rank " + std::to_string(to));
                continue;
            }
            return fail("connection to rank " + std::to_string(to) + " broke: " + std::strerror(errno));
        }
        return 0;
    }

    int recv(int from, int tag, std::vector<uint8_t>* out) {
        Peer& p = peer[from];
        const double deadline = now_s() + timeout_s;
        for (;;) {
            for (auto it = p.q.begin(); it != p.q.end(); ++it)
                if (it->tag == tag) {
                    *out = std::move(it->data);
                    p.q.erase(it);
                    return 0;
                }
            if (p.closed) return fail("rank " + std::to_string(from) + " went away");
            if (!wait_any(-1, deadline)) return fail("timed out waiting for rank " + std::to_string(from));
        }
    }

    // ---- collectives through rank 0 ------------------------------------------------------------------
    int barrier() {
        if (world == 1) return 0;
        std::vector<uint8_t> t;
        if (rank == 0) {
            for (int r = 1; r < world; ++r) if (int rc = recv(r, kTagBarrier, &t)) return rc;
            for (int r = 1; r < world; ++r) if (int rc = send(r, kTagBarrier, nullptr, 0)) return rc;
            return 0;
        }
        if (int rc = send(0, kTagBarrier, nullptr, 0)) return rc;
        return recv(0, kTagBarrier, &t);
    }

    int max_over_ranks(double* v) {
        if (world == 1) return 0;
        std::vector<uint8_t> t;
        if (rank == 0) {
            for (int r = 1; r < world; ++r) {
                if (int rc = recv(r, kTagReduce, &t)) return rc;
                double o;
                std::memcpy(&o, t.data(), 8);
                *v = std::max(*v, o);
            }
            for (int r = 1; r < world; ++r) if (int rc = send(r, kTagReduce, v, 8)) return rc;
            return 0;
        }
        if (int rc = send(0, kTagReduce, v, 8)) return rc;
        if (int rc = recv(0, kTagReduce, &t)) return rc;
        std::memcpy(v, t.data(), 8);
        return 0;
    }
};

// ---- bootstrap ------------------------------------------------------------------------------------------

extern "C" const char* pconv_net_error(const pconv_net_t* n) { return n ? n->error.c_str() : "no network"; }

extern "C" void pconv_net_close(pconv_net_t* n) {
    if (!n) return;
    for (auto& p : n->peer)
        if (p.fd >= 0) close(p.fd);
    delete n;
}

// master_fd >= 0: an already listening socket to use as rank 0's rendezvous point (pconv_native -np N binds it
// before forking, so there is no port race); otherwise rank 0 binds master_port itself.
extern "C" pconv_net_t* pconv_net_init(int rank, int world, const char* master_host, int master_port,
                                       int master_fd, double timeout_s) {
    if (rank < 0 || world < 1 || rank >= world) return nullptr;
    pconv_net* n = new pconv_net;
    n->rank = rank;
    n->world = world;
    n->timeout_s = timeout_s > 0 ? timeout_s : 60.0;
    n->peer.resize((size_t)world);
    if (world == 1) return n;
    const double deadline = now_s() + n->timeout_s;
    auto bail = [&](const char* what) -> pconv_net_t* {
        std::fprintf(stderr, "pconv net rank %d: %s: %s\n", rank, what, std::strerror(errno));
        pconv_net_close(n);
        return nullptr;
    };
    auto listen_on = [&](int port, int* chosen) -> int {
        const int fd = socket(AF_INET, SOCK_STREAM, 0);
        if (fd < 0) return -1;
        int one = 1;
        setsockopt(fd, SOL_SOCKET, SO_REUSEADDR, &one, sizeof one);
        sockaddr_in a{};
        a.sin_family = AF_INET;
        a.sin_addr.s_addr = htonl(INADDR_ANY);
        a.sin_port = htons((uint16_t)port);
        if (bind(fd, (sockaddr*)&a, sizeof a) != 0 || listen(fd, world + 8) != 0) {
            close(fd);
            return -1;
        }
        socklen_t len = sizeof a;
        getsockname(fd, (sockaddr*)&a, &len);
        if (chosen) *chosen = ntohs(a.sin_port);
        return fd;
    };
    auto connect_to = [&](uint32_t ip_be, int port) -> int {
        while (now_s() < deadline) {
            const int fd = socket(AF_INET, SOCK_STREAM, 0);
            if (fd < 0) return -1;
            sockaddr_in a{};
            a.sin_family = AF_INET;
            a.sin_addr.s_addr = ip_be;
            a.sin_port = htons((uint16_t)port);
            if (connect(fd, (sockaddr*)&a, sizeof a) == 0) return fd;
            close(fd);
            timespec nap = {0, 20 * 1000 * 1000};   // the listener may not be up yet
            nanosleep(&nap, nullptr);
        }
        return -1;
    };
    auto accept_one = [&](int lfd) -> int {
        pollfd pf{lfd, POLLIN, 0};
        const double left = deadline - now_s();
        if (left <= 0 || poll(&pf, 1, (int)(left * 1e3) + 1) <= 0) return -1;
        return accept(lfd, nullptr, nullptr);
    };

    struct Entry { uint32_t ip_be; int32_t port; };
    std::vector<Entry> table((size_t)world);
    int my_port = 0;
    int lfd = -1;
    if (rank == 0) {
        lfd = master_fd >= 0 ? master_fd : listen_on(master_port, nullptr);
        if (lfd < 0) return bail("cannot listen on the master port");
        for (int i = 1; i < world; ++i) {
            const int fd = accept_one(lfd);
            if (fd < 0) return bail("a rank did not report in");
            int32_t hello[3];
            if (!read_all(fd, hello, sizeof hello, deadline) || hello[0] != kTagHello || hello[1] <= 0 || hello[1] >= world ||
                n->peer[(size_t)hello[1]].fd >= 0) {
                close(fd);
                return bail("bad hello");
            }
            sockaddr_in a{};
            socklen_t len = sizeof a;
            getpeername(fd, (sockaddr*)&a, &len);
            table[(size_t)hello[1]] = {a.sin_addr.s_addr, hello[2]};
            n->peer[(size_t)hello[1]].fd = fd;
        }
        for (int r = 1; r < world; ++r) {
            const int32_t hdr[2] = {kTagTable, world};
            if (!write_all(n->peer[(size_t)r].fd, hdr, sizeof hdr) ||
                !write_all(n->peer[(size_t)r].fd, table.data(), sizeof(Entry) * (size_t)world))
                return bail("cannot send the rank table");
        }
        if (master_fd < 0) close(lfd);
    } else {
        lfd = listen_on(0, &my_port);
        if (lfd < 0) return bail("cannot open a listening socket");
        addrinfo hints{}, *res = nullptr;
        hints.ai_family = AF_INET;
        hints.ai_socktype = SOCK_STREAM;
        if (getaddrinfo(master_host && *master_host ? master_host : "127.0.0.1", nullptr, &hints, &res) != 0 || !res)
            return bail("cannot resolve the master host");
        const uint32_t master_ip = ((sockaddr_in*)res->ai_addr)->sin_addr.s_addr;
        freeaddrinfo(res);
        const int fd = connect_to(master_ip, master_port);
        if (fd < 0) return bail("cannot reach rank 0");
        const int32_t hello[3] = {kTagHello, rank, my_port};
        int32_t hdr[2];
        if (!write_all(fd, hello, sizeof hello) || !read_all(fd, hdr, sizeof hdr, deadline) || hdr[0] != kTagTable ||
            hdr[1] != world || !read_all(fd, table.data(), sizeof(Entry) * (size_t)world, deadline))
            return bail("rendezvous with rank 0 failed");
        n->peer[0].fd = fd;
        // full mesh among ranks 1..world-1: connect to the lower ranks, accept the higher ones
        for (int q = 1; q < rank; ++q) {
            const int c = connect_to(table[(size_t)q].ip_be, table[(size_t)q].port);
            if (c < 0) return bail("cannot reach a peer rank");
            const int32_t h2[3] = {kTagHello, rank, 0};
            if (!write_all(c, h2, sizeof h2)) return bail("peer hello");
            n->peer[(size_t)q].fd = c;
        }
        for (int q = rank + 1; q < world; ++q) {
            const int c = accept_one(lfd);
            if (c < 0) return bail("a higher rank did not connect");
            int32_t h2[3];
            if (!read_all(c, h2, sizeof h2, deadline) || h2[0] != kTagHello || h2[1] <= rank || h2[1] >= world ||
                n->peer[(size_t)h2[1]].fd >= 0) {
                close(c);
                return bail("bad peer hello");
            }
            n->peer[(size_t)h2[1]].fd = c;
        }
        close(lfd);
    }
    for (int r = 0; r < world; ++r)
        if (r != rank) {
            tune(n->peer[(size_t)r].fd);
            set_nonblock(n->peer[(size_t)r].fd, true);
        }
    if (n->barrier() != 0) return bail("first barrier");
    return n;
}

extern "C" int pconv_net_barrier(pconv_net_t* n) { return n ? n->barrier() : -EINVAL; }
extern "C" int pconv_net_max(pconv_net_t* n, double* v) { return (n && v) ? n->max_over_ranks(v) : -EINVAL; }

// ---- the iteration loop -----------------------------------------------------------------------------------

extern "C" int64_t pconv_net_tile_pitch(int64_t cols, int C, int k) { return ((cols + 2 * k) * C + 63) / 64 * 64; }

// buf_a / buf_b: two zeroed buffers of (rows + 2k) * pitch bytes, pitch = pconv_net_tile_pitch(cols, C, k); the
// tile's pixel (0,0) sits at row k, byte k*C; the input tile is in buf_a. nbr[9] as in pconv_shm_run.
// Returns 0 / 1 (result in buf_a / buf_b) or a negative errno; *exchanges counts the halo exchanges.
extern "C" int pconv_net_run(pconv_net_t* n, int64_t H, int64_t W, int C, int64_t row0, int64_t rows, int64_t col0,
                             int64_t cols, const int32_t nbr[9], uint8_t* buf_a, uint8_t* buf_b,
                             const pconv_filter_t* f, int loops, int k, int threads, int64_t* exchanges) {
    if (!n || !f || !nbr || !buf_a || !buf_b || rows <= 0 || cols <= 0 || loops < 0 || k < 1) return -EINVAL;
    const int64_t pitch = pconv_net_tile_pitch(cols, C, k);
    uint8_t* bufs[2] = {buf_a + k * pitch + (int64_t)k * C, buf_b + k * pitch + (int64_t)k * C};   // pixel (0,0)
    int cur = 0;
    int64_t nx = 0;
    const bool multi = n->world > 1;
    // tile-relative rectangle clipped to the image (outside stays zero = zero padding)
    auto conv = [&](const uint8_t* s, uint8_t* d, int64_t r0, int64_t r1, int64_t c0, int64_t c1) {
        r0 = std::max(r0, -row0);
        r1 = std::min(r1, H - row0);
        c0 = std::max(c0, -col0);
        c1 = std::min(c1, W - col0);
        if (r0 < r1 && c0 < c1) pconv_cpu_convolve(s, d, pitch, C, f, r0, r1, c0, c1, threads > 0 ? threads : 1);
    };
    // k-deep strip facing direction (dy,dx): the tile's own border (halo = false) or the halo beyond it
    auto span = [&](int d, int64_t extent, bool halo, int64_t* a, int64_t* b) {
        if (d == 0) { *a = 0; *b = extent; }
        else if (halo) { *a = d < 0 ? -k : extent; *b = d < 0 ? 0 : extent + k; }
        else { *a = d < 0 ? 0 : extent - k; *b = d < 0 ? k : extent; }
    };
    std::vector<uint8_t> pack;
    for (int done = 0; done < loops;) {
        const int nit = std::min(k, loops - done);
        uint8_t* src = bufs[cur];
        uint8_t* dst = bufs[cur ^ 1];
        if (multi) {
            // 1. send the eight border strips (the receiver's tag = the direction it sees us in)
            for (int d = 0; d < 9; ++d) {
                if (d == 4 || nbr[d] < 0) continue;
                const int dy = d / 3 - 1, dx = d % 3 - 1;
                int64_t ra, rb, ca, cb;
                span(dy, rows, false, &ra, &rb);
                span(dx, cols, false, &ca, &cb);
                const int64_t wbytes = (cb - ca) * C;
                pack.resize((size_t)((rb - ra) * wbytes));
                for (int64_t y = ra; y < rb; ++y)
                    std::memcpy(pack.data() + (y - ra) * wbytes, src + y * pitch + ca * C, (size_t)wbytes);
                if (int rc = n->send(nbr[d], 8 - d, pack.data(), (int)pack.size())) return rc;
            }
            ++nx;
            // 2. while they fly: the part of iteration 1 that reads no halo
            const int64_t ir0 = nbr[1] >= 0 ? 1 : -(int64_t)(nit - 1), ir1 = nbr[7] >= 0 ? rows - 1 : rows + nit - 1;
            const int64_t ic0 = nbr[3] >= 0 ? 1 : -(int64_t)(nit - 1), ic1 = nbr[5] >= 0 ? cols - 1 : cols + nit - 1;
            conv(src, dst, ir0, ir1, ic0, ic1);
            // 3. the halos
            std::vector<uint8_t> in;
            for (int d = 0; d < 9; ++d) {
                if (d == 4 || nbr[d] < 0) continue;
                const int dy = d / 3 - 1, dx = d % 3 - 1;
                int64_t ra, rb, ca, cb;
                span(dy, rows, true, &ra, &rb);
                span(dx, cols, true, &ca, &cb);
                const int64_t wbytes = (cb - ca) * C;
                if (int rc = n->recv(nbr[d], d, &in)) return rc;
                if ((int64_t)in.size() != (rb - ra) * wbytes) return n->fail("halo strip of unexpected size");
                for (int64_t y = ra; y < rb; ++y)
                    std::memcpy(src + y * pitch + ca * C, in.data() + (y - ra) * wbytes, (size_t)wbytes);
            }
            // 4. the frame of iteration 1 around what step 2 covered
            const int64_t m = nit - 1;
            conv(src, dst, -m, ir0, -m, cols + m);            // top band
            conv(src, dst, ir1, rows + m, -m, cols + m);      // bottom band
            conv(src, dst, ir0, ir1, -m, ic0);                // left
            conv(src, dst, ir0, ir1, ic1, cols + m);          // right
            cur ^= 1;
            // 5. iterations 2..nit over the shrinking ring
            for (int j = 2; j <= nit; ++j) {
                const int64_t mm = nit - j;
                conv(bufs[cur], bufs[cur ^ 1], -mm, rows + mm, -mm, cols + mm);
                cur ^= 1;
            }
        } else {
            for (int j = 1; j <= nit; ++j) {
                conv(bufs[cur], bufs[cur ^ 1], 0, rows, 0, cols);
                cur ^= 1;
            }
        }
        done += nit;
    }
    if (exchanges) *exchanges = nx;
    return cur;
}
