/* Implementation of the MPI subset declared in mpi.h (see the comment there).
 *
 * Process model: mpirun_shim forks one process per rank and hands every rank one connected AF_UNIX
 * stream socket per peer (PCONV_MPI_FDS). Without that environment the program runs as a singleton
 * (size 1), so `./mpi_conv ...` works directly for one process.
 *
 * Sends are eager: the (packed) message is written to the peer's socket at MPI_Isend/MPI_Send time.
 * Whenever a write or a wait would block, every peer's socket is drained into a per-peer queue of
 * unexpected messages first, so two ranks that send to each other before receiving cannot deadlock
 * regardless of message size. Receives match by (source, tag) in posting order; a stream socket is
 * FIFO, which gives MPI's non-overtaking guarantee.
 */
#define _GNU_SOURCE
#include "mpi.h"

#include <errno.h>
#include <fcntl.h>
#include <poll.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <unistd.h>

#define TAG_BCAST (-100)
#define TAG_BARRIER (-101)
#define MAX_TYPES 64
#define FIRST_DERIVED 16

typedef struct { int used, nblocks, blockbytes, stridebytes; } dtype_t;

typedef struct msg { struct msg* next; int tag, nbytes; unsigned char data[]; } msg_t;

struct pconv_mpi_request {
    int is_recv, done, peer, tag, count;
    MPI_Datatype type;
    void* buf;
    struct pconv_mpi_request* next;   /* posted-receive list of the peer */
};

struct pconv_mpi_file { int fd; };

typedef struct {
    int fd, closed;         /* closed: the peer has shut its end (it finalized or died) */
    unsigned char* rx;      /* bytes read but not yet a complete message */
    size_t rx_len, rx_cap;
    msg_t *q_head, *q_tail; /* complete messages nobody has received yet */
    struct pconv_mpi_request *posted_head, *posted_tail;
} peer_t;

static int g_rank = 0, g_size = 1, g_init = 0;
static peer_t* g_peer = NULL;
static dtype_t g_types[MAX_TYPES];

static void die(const char* what) {
    fprintf(stderr, "[mpi_shim rank %d] %s: %s\n", g_rank, what, strerror(errno));
    _exit(70);
}

/* ---- datatypes --------------------------------------------------------------------------------- */

static int elem_size(MPI_Datatype t) {
    switch (t) {
        case MPI_BYTE: case MPI_CHAR: return 1;
        case MPI_INT: case MPI_FLOAT: return 4;
        case MPI_DOUBLE: return 8;
        default: return 0;
    }
}

static dtype_t describe(MPI_Datatype t) {
    dtype_t d = {1, 1, 0, 0};
    if (t >= FIRST_DERIVED && t < MAX_TYPES && g_types[t].used) return g_types[t];
    d.blockbytes = d.stridebytes = elem_size(t);
    if (d.blockbytes == 0) { errno = EINVAL; die("unknown datatype"); }
    return d;
}

static size_t type_size(MPI_Datatype t) { dtype_t d = describe(t); return (size_t)d.nblocks * d.blockbytes; }
static size_t type_extent(MPI_Datatype t) {
    dtype_t d = describe(t);
    return (size_t)(d.nblocks - 1) * d.stridebytes + d.blockbytes;
}

static void pack(const void* buf, int count, MPI_Datatype t, unsigned char* out) {
    dtype_t d = describe(t);
    const size_t ext = type_extent(t);
    for (int c = 0; c < count; ++c) {
        const unsigned char* base = (const unsigned char*)buf + (size_t)c * ext;
        for (int b = 0; b < d.nblocks; ++b, out += d.blockbytes)
            memcpy(out, base + (size_t)b * d.stridebytes, (size_t)d.blockbytes);
    }
}

static void unpack(void* buf, int count, MPI_Datatype t, const unsigned char* in, size_t avail) {
    dtype_t d = describe(t);
    const size_t ext = type_extent(t);
    for (int c = 0; c < count; ++c) {
        unsigned char* base = (unsigned char*)buf + (size_t)c * ext;
        for (int b = 0; b < d.nblocks; ++b) {
            size_t n = (size_t)d.blockbytes < avail ? (size_t)d.blockbytes : avail;
            memcpy(base + (size_t)b * d.stridebytes, in, n);
            in += n;
            avail -= n;
            if (avail == 0) return;
        }
    }
}

static int new_type(int nblocks, int blockbytes, int stridebytes) {
    for (int i = FIRST_DERIVED; i < MAX_TYPES; ++i)
        if (!g_types[i].used) {
            g_types[i] = (dtype_t){1, nblocks, blockbytes, stridebytes};
            return i;
        }
    errno = ENOMEM;
    die("too many datatypes");
    return -1;
}

int MPI_Type_vector(int count, int blocklength, int stride, MPI_Datatype oldtype, MPI_Datatype* newtype) {
    const int es = (int)type_size(oldtype);   /* only contiguous old types are needed here */
    *newtype = new_type(count, blocklength * es, stride * es);
    return MPI_SUCCESS;
}

int MPI_Type_contiguous(int count, MPI_Datatype oldtype, MPI_Datatype* newtype) {
    const int es = (int)type_size(oldtype);
    *newtype = new_type(1, count * es, count * es);
    return MPI_SUCCESS;
}

int MPI_Type_commit(MPI_Datatype* type) { (void)type; return MPI_SUCCESS; }

int MPI_Type_free(MPI_Datatype* type) {
    if (*type >= FIRST_DERIVED && *type < MAX_TYPES) g_types[*type].used = 0;
    *type = 0;
    return MPI_SUCCESS;
}

/* ---- transport --------------------------------------------------------------------------------- */

static void parse_rx(peer_t* p) {
    size_t off = 0;
    while (p->rx_len - off >= 2 * sizeof(int)) {
        int hdr[2];
        memcpy(hdr, p->rx + off, sizeof hdr);
        const size_t need = sizeof hdr + (size_t)hdr[1];
        if (p->rx_len - off < need) break;
        msg_t* m = (msg_t*)malloc(sizeof(msg_t) + (size_t)hdr[1]);
        if (!m) die("malloc");
        m->next = NULL;
        m->tag = hdr[0];
        m->nbytes = hdr[1];
        memcpy(m->data, p->rx + off + sizeof hdr, (size_t)hdr[1]);
        if (p->q_tail) p->q_tail->next = m; else p->q_head = m;
        p->q_tail = m;
        off += need;
    }
    if (off) {
        memmove(p->rx, p->rx + off, p->rx_len - off);
        p->rx_len -= off;
    }
}

/* read whatever the socket holds right now; returns bytes read */
static size_t drain(peer_t* p) {
    size_t total = 0;
    for (;;) {
        if (p->rx_cap - p->rx_len < 65536) {
            p->rx_cap = p->rx_cap ? p->rx_cap * 2 : (1 << 17);
            p->rx = (unsigned char*)realloc(p->rx, p->rx_cap);
            if (!p->rx) die("realloc");
        }
        ssize_t n = read(p->fd, p->rx + p->rx_len, p->rx_cap - p->rx_len);
        if (n > 0) { p->rx_len += (size_t)n; total += (size_t)n; continue; }
        if (n == 0) { p->closed = 1; break; }   /* fatal only if a message from it is still awaited */
        if (errno == EAGAIN || errno == EWOULDBLOCK) break;
        if (errno == EINTR) continue;
        die("read");
    }
    if (total) parse_rx(p);
    return total;
}

/* block until some peer has data (or `out_fd` is writable), then drain every readable peer */
static void wait_any(int out_fd) {
    struct pollfd pf[g_size + 1];
    int n = 0;
    for (int r = 0; r < g_size; ++r)
        if (r != g_rank && !g_peer[r].closed) pf[n++] = (struct pollfd){g_peer[r].fd, POLLIN, 0};
    if (out_fd >= 0) pf[n++] = (struct pollfd){out_fd, POLLOUT, 0};
    while (poll(pf, (nfds_t)n, -1) < 0)
        if (errno != EINTR) die("poll");
    for (int k = 0; k < n; ++k) {
        if (pf[k].fd == out_fd && (pf[k].events & POLLOUT)) continue;
        if (pf[k].revents & (POLLIN | POLLHUP | POLLERR))
            for (int r = 0; r < g_size; ++r)
                if (r != g_rank && g_peer[r].fd == pf[k].fd) drain(&g_peer[r]);
    }
}

static void send_packed(int dest, int tag, const unsigned char* payload, int nbytes) {
    peer_t* p = &g_peer[dest];
    const size_t total = 2 * sizeof(int) + (size_t)nbytes;
    unsigned char* frame = (unsigned char*)malloc(total);
    if (!frame) die("malloc");
    int hdr[2] = {tag, nbytes};
    memcpy(frame, hdr, sizeof hdr);
    if (nbytes) memcpy(frame + sizeof hdr, payload, (size_t)nbytes);
    size_t off = 0;
    while (off < total) {
        ssize_t n = write(p->fd, frame + off, total - off);
        if (n > 0) { off += (size_t)n; continue; }
        if (n < 0 && errno == EINTR) continue;
        if (n < 0 && (errno == EAGAIN || errno == EWOULDBLOCK)) { wait_any(p->fd); continue; }
        die("write");
    }
    free(frame);
}

static int try_match(peer_t* p, struct pconv_mpi_request* r) {
    msg_t* prev = NULL;
    for (msg_t* m = p->q_head; m; prev = m, m = m->next) {
        if (m->tag != r->tag) continue;
        unpack(r->buf, r->count, r->type, m->data, (size_t)m->nbytes);
        if (prev) prev->next = m->next; else p->q_head = m->next;
        if (p->q_tail == m) p->q_tail = prev;
        free(m);
        r->done = 1;
        return 1;
    }
    return 0;
}

/* give queued messages to the posted receives of this peer, oldest receive first */
static void match_posted(peer_t* p) {
    struct pconv_mpi_request* prev = NULL;
    for (struct pconv_mpi_request* r = p->posted_head; r;) {
        struct pconv_mpi_request* next = r->next;
        if (try_match(p, r)) {
            if (prev) prev->next = next; else p->posted_head = next;
            if (p->posted_tail == r) p->posted_tail = prev;
            r->next = NULL;
        } else {
            prev = r;
        }
        r = next;
    }
}

static struct pconv_mpi_request* post_recv(void* buf, int count, MPI_Datatype type, int source, int tag) {
    struct pconv_mpi_request* r = (struct pconv_mpi_request*)calloc(1, sizeof *r);
    if (!r) die("calloc");
    r->is_recv = 1;
    r->peer = source;
    r->tag = tag;
    r->count = count;
    r->type = type;
    r->buf = buf;
    peer_t* p = &g_peer[source];
    if (p->posted_tail) p->posted_tail->next = r; else p->posted_head = r;
    p->posted_tail = r;
    match_posted(p);
    return r;
}

static void complete(struct pconv_mpi_request* r) {
    if (!r->is_recv) return;
    peer_t* p = &g_peer[r->peer];
    while (!r->done) {
        match_posted(p);
        if (r->done) break;
        if (p->closed) {
            fprintf(stderr, "[mpi_shim rank %d] rank %d exited while a message from it was awaited\n", g_rank, r->peer);
            _exit(71);
        }
        wait_any(-1);
    }
}

static void check_peer(int r, const char* what) {
    if (r < 0 || r >= g_size || r == g_rank) {
        fprintf(stderr, "[mpi_shim rank %d] %s: invalid peer %d\n", g_rank, what, r);
        _exit(72);
    }
}

/* ---- public API -------------------------------------------------------------------------------- */

int MPI_Init(int* argc, char*** argv) {
    (void)argc; (void)argv;
    const char* rk = getenv("PCONV_MPI_RANK");
    const char* sz = getenv("PCONV_MPI_SIZE");
    const char* fds = getenv("PCONV_MPI_FDS");
    if (rk && sz && fds) {
        g_rank = atoi(rk);
        g_size = atoi(sz);
    }
    g_peer = (peer_t*)calloc((size_t)g_size, sizeof(peer_t));
    if (!g_peer) die("calloc");
    for (int r = 0; r < g_size; ++r) g_peer[r].fd = -1;
    if (g_size > 1) {
        const char* s = fds;
        for (int r = 0; r < g_size; ++r) {
            char* end;
            g_peer[r].fd = (int)strtol(s, &end, 10);
            s = (*end == ',') ? end + 1 : end;
            if (r != g_rank) {
                int fl = fcntl(g_peer[r].fd, F_GETFL);
                if (fl < 0 || fcntl(g_peer[r].fd, F_SETFL, fl | O_NONBLOCK) < 0) die("fcntl");
            }
        }
    }
    g_init = 1;
    return MPI_SUCCESS;
}

int MPI_Finalize(void) {
    if (g_size > 1) MPI_Barrier(MPI_COMM_WORLD);   /* nobody closes a socket a peer still writes to */
    for (int r = 0; r < g_size; ++r)
        if (g_peer[r].fd >= 0 && r != g_rank) close(g_peer[r].fd);
    g_init = 0;
    return MPI_SUCCESS;
}

int MPI_Abort(MPI_Comm comm, int errorcode) {
    (void)comm;
    fflush(NULL);
    _exit(errorcode ? errorcode : 1);   /* mpirun_shim stops the other ranks when one exits non-zero */
}

int MPI_Comm_size(MPI_Comm comm, int* size) { (void)comm; *size = g_size; return MPI_SUCCESS; }
int MPI_Comm_rank(MPI_Comm comm, int* rank) { (void)comm; *rank = g_rank; return MPI_SUCCESS; }

double MPI_Wtime(void) {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

int MPI_Send(const void* buf, int count, MPI_Datatype type, int dest, int tag, MPI_Comm comm) {
    (void)comm;
    if (dest == MPI_PROC_NULL) return MPI_SUCCESS;
    check_peer(dest, "send");
    const size_t n = type_size(type) * (size_t)count;
    unsigned char* tmp = (unsigned char*)malloc(n ? n : 1);
    if (!tmp) die("malloc");
    pack(buf, count, type, tmp);
    send_packed(dest, tag, tmp, (int)n);
    free(tmp);
    return MPI_SUCCESS;
}

int MPI_Isend(const void* buf, int count, MPI_Datatype type, int dest, int tag, MPI_Comm comm, MPI_Request* req) {
    MPI_Send(buf, count, type, dest, tag, comm);   /* eager: complete once written to the socket */
    struct pconv_mpi_request* r = (struct pconv_mpi_request*)calloc(1, sizeof *r);
    if (!r) die("calloc");
    r->done = 1;
    *req = r;
    return MPI_SUCCESS;
}

int MPI_Irecv(void* buf, int count, MPI_Datatype type, int source, int tag, MPI_Comm comm, MPI_Request* req) {
    (void)comm;
    if (source == MPI_PROC_NULL) {
        struct pconv_mpi_request* r = (struct pconv_mpi_request*)calloc(1, sizeof *r);
        if (!r) die("calloc");
        r->done = 1;
        *req = r;
        return MPI_SUCCESS;
    }
    check_peer(source, "recv");
    *req = post_recv(buf, count, type, source, tag);
    return MPI_SUCCESS;
}

int MPI_Wait(MPI_Request* req, MPI_Status* status) {
    if (!req || *req == MPI_REQUEST_NULL) return MPI_SUCCESS;
    struct pconv_mpi_request* r = *req;
    complete(r);
    if (status) {
        status->MPI_SOURCE = r->peer;
        status->MPI_TAG = r->tag;
        status->MPI_ERROR = MPI_SUCCESS;
    }
    free(r);
    *req = MPI_REQUEST_NULL;
    return MPI_SUCCESS;
}

int MPI_Recv(void* buf, int count, MPI_Datatype type, int source, int tag, MPI_Comm comm, MPI_Status* status) {
    MPI_Request r;
    MPI_Irecv(buf, count, type, source, tag, comm, &r);
    return MPI_Wait(&r, status);
}

int MPI_Bcast(void* buf, int count, MPI_Datatype type, int root, MPI_Comm comm) {
    if (g_size == 1) return MPI_SUCCESS;
    if (g_rank == root) {
        for (int r = 0; r < g_size; ++r)
            if (r != root) MPI_Send(buf, count, type, r, TAG_BCAST, comm);
    } else {
        MPI_Recv(buf, count, type, root, TAG_BCAST, comm, MPI_STATUS_IGNORE);
    }
    return MPI_SUCCESS;
}

int MPI_Barrier(MPI_Comm comm) {
    if (g_size == 1) return MPI_SUCCESS;
    char token = 0;
    if (g_rank == 0) {
        for (int r = 1; r < g_size; ++r) MPI_Recv(&token, 1, MPI_BYTE, r, TAG_BARRIER, comm, MPI_STATUS_IGNORE);
        for (int r = 1; r < g_size; ++r) MPI_Send(&token, 1, MPI_BYTE, r, TAG_BARRIER, comm);
    } else {
        MPI_Send(&token, 1, MPI_BYTE, 0, TAG_BARRIER, comm);
        MPI_Recv(&token, 1, MPI_BYTE, 0, TAG_BARRIER, comm, MPI_STATUS_IGNORE);
    }
    return MPI_SUCCESS;
}

/* ---- files: every rank opens the file itself (the reference only uses independent seek/read/write) */

int MPI_File_open(MPI_Comm comm, const char* filename, int amode, MPI_Info info, MPI_File* fh) {
    (void)comm; (void)info;
    int flags = (amode & MPI_MODE_RDWR) ? O_RDWR : (amode & MPI_MODE_WRONLY) ? O_WRONLY : O_RDONLY;
    if (amode & MPI_MODE_CREATE) flags |= O_CREAT;
    int fd = open(filename, flags, 0644);
    if (fd < 0) {
        fprintf(stderr, "[mpi_shim rank %d] cannot open %s: %s\n", g_rank, filename, strerror(errno));
        MPI_Abort(comm, 2);
    }
    *fh = (MPI_File)malloc(sizeof **fh);
    if (!*fh) die("malloc");
    (*fh)->fd = fd;
    return MPI_SUCCESS;
}

int MPI_File_close(MPI_File* fh) {
    if (fh && *fh) {
        close((*fh)->fd);
        free(*fh);
        *fh = NULL;
    }
    return MPI_SUCCESS;
}

int MPI_File_seek(MPI_File fh, MPI_Offset offset, int whence) {
    const int w = whence == MPI_SEEK_SET ? SEEK_SET : whence == MPI_SEEK_CUR ? SEEK_CUR : SEEK_END;
    if (lseek(fh->fd, (off_t)offset, w) == (off_t)-1) die("lseek");
    return MPI_SUCCESS;
}

int MPI_File_read(MPI_File fh, void* buf, int count, MPI_Datatype type, MPI_Status* status) {
    const size_t want = type_size(type) * (size_t)count;
    size_t got = 0;
    while (got < want) {
        ssize_t n = read(fh->fd, (char*)buf + got, want - got);
        if (n > 0) { got += (size_t)n; continue; }
        if (n == 0) break;   /* end of file: a short read, like MPI */
        if (errno == EINTR) continue;
        die("file read");
    }
    if (status) status->count_ = (int)got;
    return MPI_SUCCESS;
}

int MPI_File_write(MPI_File fh, const void* buf, int count, MPI_Datatype type, MPI_Status* status) {
    const size_t want = type_size(type) * (size_t)count;
    size_t put = 0;
    while (put < want) {
        ssize_t n = write(fh->fd, (const char*)buf + put, want - put);
        if (n > 0) { put += (size_t)n; continue; }
        if (n < 0 && errno == EINTR) continue;
        die("file write");
    }
    if (status) status->count_ = (int)put;
    return MPI_SUCCESS;
}
