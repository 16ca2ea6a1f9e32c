/* mpirun_shim -np N prog [args...]: start N ranks of a program linked against mpi_shim.c on this node.
 *
 * Creates one AF_UNIX stream socket pair per pair of ranks, forks N children, gives child r the
 * descriptors of its N-1 channels through PCONV_MPI_FDS and execs the program. If a rank exits
 * non-zero (MPI_Abort, crash), the remaining ranks are stopped by PID and the launcher fails.
 */
#define _GNU_SOURCE
#include <errno.h>
#include <signal.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/resource.h>
#include <sys/socket.h>
#include <sys/wait.h>
#include <unistd.h>

static void usage(void) {
    fprintf(stderr, "usage: mpirun_shim -np N prog [args...]\n");
    exit(2);
}

int main(int argc, char** argv) {
    int np = 0, ai = 1;
    while (ai < argc && argv[ai][0] == '-') {
        if ((!strcmp(argv[ai], "-np") || !strcmp(argv[ai], "-n")) && ai + 1 < argc) {
            np = atoi(argv[ai + 1]);
            ai += 2;
        } else if (!strcmp(argv[ai], "--oversubscribe")) {
            ai += 1;   /* accepted for command-line compatibility; ranks are plain processes anyway */
        } else {
            usage();
        }
    }
    if (np < 1 || np > 128 || ai >= argc) usage();

    struct rlimit rl;
    if (getrlimit(RLIMIT_NOFILE, &rl) == 0 && rl.rlim_cur < rl.rlim_max) {
        rl.rlim_cur = rl.rlim_max;
        setrlimit(RLIMIT_NOFILE, &rl);
    }

    int* fd = (int*)malloc(sizeof(int) * (size_t)np * (size_t)np);   /* fd[r*np + p]: rank r's end of channel r<->p */
    if (!fd) return 1;
    for (int i = 0; i < np * np; ++i) fd[i] = -1;
    for (int a = 0; a < np; ++a)
        for (int b = a + 1; b < np; ++b) {
            int sv[2];
            if (socketpair(AF_UNIX, SOCK_STREAM, 0, sv) != 0) {
                perror("mpirun_shim: socketpair");
                return 1;
            }
            int sz = 1 << 20;
            setsockopt(sv[0], SOL_SOCKET, SO_SNDBUF, &sz, sizeof sz);
            setsockopt(sv[1], SOL_SOCKET, SO_SNDBUF, &sz, sizeof sz);
            fd[a * np + b] = sv[0];
            fd[b * np + a] = sv[1];
        }

    pid_t* pid = (pid_t*)calloc((size_t)np, sizeof(pid_t));
    if (!pid) return 1;
    for (int r = 0; r < np; ++r) {
        pid[r] = fork();
        if (pid[r] < 0) {
            perror("mpirun_shim: fork");
            for (int k = 0; k < r; ++k) kill(pid[k], SIGTERM);
            return 1;
        }
        if (pid[r] == 0) {
            char list[16 * 128 + 1], num[32];
            list[0] = 0;
            for (int p = 0; p < np; ++p) {
                snprintf(num, sizeof num, p ? ",%d" : "%d", fd[r * np + p]);
                strcat(list, num);
            }
            for (int q = 0; q < np; ++q)      /* keep only this rank's ends */
                for (int p = 0; p < np; ++p)
                    if (q != r && fd[q * np + p] >= 0) close(fd[q * np + p]);
            snprintf(num, sizeof num, "%d", r);
            setenv("PCONV_MPI_RANK", num, 1);
            snprintf(num, sizeof num, "%d", np);
            setenv("PCONV_MPI_SIZE", num, 1);
            setenv("PCONV_MPI_FDS", list, 1);
            execvp(argv[ai], argv + ai);
            fprintf(stderr, "mpirun_shim: cannot exec %s: %s\n", argv[ai], strerror(errno));
            _exit(127);
        }
    }
    for (int i = 0; i < np * np; ++i)
        if (fd[i] >= 0) close(fd[i]);

    int failed = 0, left = np;
    while (left > 0) {
        int st;
        pid_t w = wait(&st);
        if (w < 0) {
            if (errno == EINTR) continue;
            break;
        }
        --left;
        int bad = !(WIFEXITED(st) && WEXITSTATUS(st) == 0);
        for (int r = 0; r < np; ++r)
            if (pid[r] == w) pid[r] = 0;
        if (bad) {
            /* 71 = "a peer vanished under me" (mpi_shim.c): a consequence, not the cause. The status reported
             * is the first real failure; ranks stopped by our own SIGTERM never count. */
            const int code = WIFEXITED(st) ? WEXITSTATUS(st) : 128 + WTERMSIG(st);
            const int ours = WIFSIGNALED(st) && WTERMSIG(st) == SIGTERM && failed;
            if (!failed || (failed == 71 && !ours && code != 71)) {
                if (!failed)
                    for (int r = 0; r < np; ++r)
                        if (pid[r] > 0) kill(pid[r], SIGTERM);
                failed = code;
            }
        }
    }
    return failed;
}
