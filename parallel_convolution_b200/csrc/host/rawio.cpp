// Raw headerless image I/O with 64-bit offsets, rectangle-at-a-time, plus the synthetic image
// generator. Replaces the reference's per-row MPI_File_seek/read/write loops
// (mpi/mpi_conv.c:133-147, 253-271) and its POSIX read_all/write_all (cuda/funcs.c:31-45).
// All offsets are int64_t: the reference's `int filesize = width*3*height` overflows at
// 32768x32768 RGB (defect D6).
#include <errno.h>
#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>

#include <cstdint>
#include <cstring>

#include "pconv_host.h"

namespace {

int pread_full(int fd, uint8_t* buf, int64_t n, int64_t off) {
    while (n > 0) {
        const ssize_t got = ::pread(fd, buf, (size_t)n, (off_t)off);
        if (got < 0) {
            if (errno == EINTR) continue;
            return -errno;
        }
        if (got == 0) return -EIO;  // short file: fail instead of spinning
        buf += got;
        off += got;
        n -= got;
    }
    return 0;
}

int pwrite_full(int fd, const uint8_t* buf, int64_t n, int64_t off) {
    while (n > 0) {
        const ssize_t put = ::pwrite(fd, buf, (size_t)n, (off_t)off);
        if (put < 0) {
            if (errno == EINTR) continue;
            return -errno;
        }
        buf += put;
        off += put;
        n -= put;
    }
    return 0;
}

inline uint64_t splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

}  // namespace

extern "C" int64_t pconv_file_size(const char* path) {
    struct stat st;
    if (::stat(path, &st) != 0) return -(int64_t)errno;
    return (int64_t)st.st_size;
}

extern "C" int pconv_read_rect(const char* path, int64_t row_bytes, int64_t row0, int64_t nrows,
                               int64_t byte0, int64_t nbytes, uint8_t* dst, int64_t dst_pitch) {
    const int fd = ::open(path, O_RDONLY);
    if (fd < 0) return -errno;
    int rc = 0;
    if (byte0 == 0 && nbytes == row_bytes && dst_pitch == row_bytes) {
        rc = pread_full(fd, dst, nrows * row_bytes, row0 * row_bytes);  // one contiguous read
    } else {
        for (int64_t r = 0; r < nrows && rc == 0; ++r)
            rc = pread_full(fd, dst + r * dst_pitch, nbytes, (row0 + r) * row_bytes + byte0);
    }
    ::close(fd);
    return rc;
}

extern "C" int pconv_create_file(const char* path, int64_t total_bytes) {
    // O_TRUNC: the reference opens without it (defect D8) and can leave stale tails behind.
    const int fd = ::open(path, O_CREAT | O_WRONLY | O_TRUNC, 0644);
    if (fd < 0) return -errno;
    int rc = 0;
    if (::ftruncate(fd, (off_t)total_bytes) != 0) rc = -errno;
    ::close(fd);
    return rc;
}

extern "C" int pconv_write_rect(const char* path, int64_t row_bytes, int64_t row0, int64_t nrows,
                                int64_t byte0, int64_t nbytes, const uint8_t* src,
                                int64_t src_pitch) {
    const int fd = ::open(path, O_WRONLY);
    if (fd < 0) return -errno;
    int rc = 0;
    if (byte0 == 0 && nbytes == row_bytes && src_pitch == row_bytes) {
        rc = pwrite_full(fd, src, nrows * row_bytes, row0 * row_bytes);
    } else {
        for (int64_t r = 0; r < nrows && rc == 0; ++r)
            rc = pwrite_full(fd, src + r * src_pitch, nbytes, (row0 + r) * row_bytes + byte0);
    }
    ::close(fd);
    return rc;
}

extern "C" void pconv_synth_rect(uint64_t seed, int64_t row_bytes, int64_t row0, int64_t nrows,
                                 int64_t byte0, int64_t nbytes, uint8_t* dst, int64_t dst_pitch) {
    // byte i of the image = byte (i & 7) of splitmix64(seed ^ (i >> 3)); the CUDA generator
    // (csrc/cuda/runtime.cu) computes the same function.
    for (int64_t r = 0; r < nrows; ++r) {
        uint8_t* out = dst + r * dst_pitch;
        const int64_t base = (row0 + r) * row_bytes + byte0;
        for (int64_t i = 0; i < nbytes; ++i) {
            const uint64_t g = (uint64_t)(base + i);
            const uint64_t word = splitmix64(seed ^ (g >> 3));
            out[i] = (uint8_t)(word >> (8 * (g & 7)));
        }
    }
}

extern "C" const char* pconv_host_version(void) { return "pconv-host 0.1"; }
