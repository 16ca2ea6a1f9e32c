/* Self-test of the MPI stand-in: run as `mpirun_shim -np N ./selftest` (N >= 2) or alone (N = 1).
 * Exercises what the reference programs rely on: broadcast, barrier, strided datatypes, head-to-head
 * non-blocking exchanges (also far larger than a socket buffer), waits in a different order than the
 * posts, message order between one pair, a max-reduction by Send/Recv, independent file access.
 * Prints "selftest ok <ranks>" from rank 0 and exits 0, or aborts with a message.
 */
#include <mpi.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

static int rank, size;

#define CHECK(cond, what)                                                                  \
    do {                                                                                   \
        if (!(cond)) {                                                                     \
            fprintf(stderr, "selftest rank %d: %s failed (line %d)\n", rank, what, __LINE__); \
            MPI_Abort(MPI_COMM_WORLD, 3);                                                  \
        }                                                                                  \
    } while (0)

int main(int argc, char** argv) {
    MPI_Init(&argc, &argv);
    MPI_Comm_size(MPI_COMM_WORLD, &size);
    MPI_Comm_rank(MPI_COMM_WORLD, &rank);
    MPI_Status st;

    /* broadcast */
    int v[3] = {0, 0, 0};
    if (rank == 0) { v[0] = 7; v[1] = -1; v[2] = 1 << 20; }
    for (int i = 0; i < 3; ++i) MPI_Bcast(&v[i], 1, MPI_INT, 0, MPI_COMM_WORLD);
    CHECK(v[0] == 7 && v[1] == -1 && v[2] == (1 << 20), "bcast");
    MPI_Barrier(MPI_COMM_WORLD);

    if (size > 1) {
        const int right = (rank + 1) % size, left = (rank + size - 1) % size;

        /* strided column out of an 8x10 byte tile into the neighbour's column, head to head */
        enum { R = 8, P = 10 };
        unsigned char tile[R * P], got[R * P];
        for (int i = 0; i < R * P; ++i) tile[i] = (unsigned char)(rank * 16 + i);
        memset(got, 0xEE, sizeof got);
        MPI_Datatype col, row;
        MPI_Type_vector(R, 3, P, MPI_BYTE, &col);
        MPI_Type_commit(&col);
        MPI_Type_contiguous(P - 2, MPI_BYTE, &row);
        MPI_Type_commit(&row);
        MPI_Request s1, s2, r1, r2;
        MPI_Isend(tile + 1, 1, col, right, 0, MPI_COMM_WORLD, &s1);
        MPI_Irecv(got + 4, 1, col, left, 0, MPI_COMM_WORLD, &r1);
        MPI_Isend(tile + P + 1, 1, row, right, 0, MPI_COMM_WORLD, &s2);   /* second message, same pair, same tag */
        MPI_Irecv(got + 7 * P + 1, 1, row, left, 0, MPI_COMM_WORLD, &r2);
        MPI_Wait(&r2, &st);   /* waited in the opposite order of posting: matching must follow the posts */
        MPI_Wait(&r1, &st);
        CHECK(st.MPI_SOURCE == left && st.MPI_TAG == 0, "status");
        MPI_Wait(&s1, &st);
        MPI_Wait(&s2, &st);
        for (int i = 0; i < R; ++i)
            for (int j = 0; j < P; ++j) {
                unsigned char want = 0xEE;
                if (j >= 4 && j < 7) want = (unsigned char)(left * 16 + i * P + (j - 4) + 1);
                if (i == 7 && j >= 1 && j < 1 + P - 2 && !(j >= 4 && j < 7)) want = (unsigned char)(left * 16 + P + 1 + (j - 1));
                if (i == 7 && j >= 4 && j < 7) want = (unsigned char)(left * 16 + P + 1 + (j - 1));   /* the row landed last */
                CHECK(got[i * P + j] == want, "strided exchange");
            }
        MPI_Type_free(&col);
        MPI_Type_free(&row);

        /* both directions at once, far larger than any socket buffer: must not deadlock */
        const int big = 8 << 20;
        unsigned char *out = (unsigned char*)malloc((size_t)big), *in = (unsigned char*)malloc((size_t)big);
        CHECK(out && in, "malloc");
        for (int i = 0; i < big; ++i) out[i] = (unsigned char)(i * 31 + rank);
        MPI_Request sb, rb;
        MPI_Isend(out, big, MPI_BYTE, right, 5, MPI_COMM_WORLD, &sb);
        MPI_Irecv(in, big, MPI_BYTE, left, 5, MPI_COMM_WORLD, &rb);
        MPI_Wait(&rb, &st);
        MPI_Wait(&sb, &st);
        for (int i = 0; i < big; i += 4099) CHECK(in[i] == (unsigned char)(i * 31 + left), "large exchange");
        free(out);
        free(in);

        /* tags are matched, not just arrival order */
        int a = 100 + rank, b = 200 + rank, ga = 0, gb = 0;
        MPI_Send(&a, 1, MPI_INT, right, 11, MPI_COMM_WORLD);
        MPI_Send(&b, 1, MPI_INT, right, 12, MPI_COMM_WORLD);
        MPI_Recv(&gb, 1, MPI_INT, left, 12, MPI_COMM_WORLD, &st);
        MPI_Recv(&ga, 1, MPI_INT, left, 11, MPI_COMM_WORLD, &st);
        CHECK(ga == 100 + left && gb == 200 + left, "tag matching");
    }

    /* max over ranks the way the reference does it */
    double mine = 1.0 + rank, top = mine;
    if (rank != 0) {
        MPI_Send(&mine, 1, MPI_DOUBLE, 0, 0, MPI_COMM_WORLD);
    } else {
        for (int i = 1; i < size; ++i) {
            double other;
            MPI_Recv(&other, 1, MPI_DOUBLE, i, 0, MPI_COMM_WORLD, &st);
            if (other > top) top = other;
        }
        CHECK(top == (double)size, "max over ranks");
    }

    /* every rank writes its own slice of one file, then reads a neighbour's */
    char path[256];
    snprintf(path, sizeof path, "/tmp/pconv_mpi_selftest_%d.bin", v[2] + size);
    if (rank == 0) unlink(path);
    MPI_Barrier(MPI_COMM_WORLD);
    MPI_File fh;
    unsigned char slice[64];
    memset(slice, rank + 1, sizeof slice);
    MPI_File_open(MPI_COMM_WORLD, path, MPI_MODE_CREATE | MPI_MODE_WRONLY, MPI_INFO_NULL, &fh);
    MPI_File_seek(fh, (MPI_Offset)rank * 64, MPI_SEEK_SET);
    MPI_File_write(fh, slice, 64, MPI_BYTE, MPI_STATUS_IGNORE);
    MPI_File_close(&fh);
    MPI_Barrier(MPI_COMM_WORLD);
    const int other = (rank + 1) % size;
    MPI_File_open(MPI_COMM_WORLD, path, MPI_MODE_RDONLY, MPI_INFO_NULL, &fh);
    MPI_File_seek(fh, (MPI_Offset)other * 64, MPI_SEEK_SET);
    MPI_File_read(fh, slice, 64, MPI_BYTE, &st);
    MPI_File_close(&fh);
    for (int i = 0; i < 64; ++i) CHECK(slice[i] == other + 1, "file slices");
    MPI_Barrier(MPI_COMM_WORLD);
    if (rank == 0) unlink(path);

    double t0 = MPI_Wtime();
    CHECK(MPI_Wtime() >= t0, "wtime");
    if (rank == 0) printf("selftest ok %d\n", size);
    MPI_Finalize();
    return 0;
}
