/* minimpirun -- launcher of the minimpi multi-process back-end (minimpi_shm.c).
 *
 *     minimpirun -n N [--ring-kb K] program [args...]
 *
 * Creates the shared segment (N*N rings of K KiB, default 1024 capped so that the
 * whole segment stays below 512 MiB), starts N copies of `program` with
 * MINIMPI_SHM / MINIMPI_RANK / MINIMPI_SIZE (and RANK / LOCAL_RANK / WORLD_SIZE for
 * code that selects its GPU from them), waits for all of them, and removes the
 * segment.  If a rank dies or aborts, the abort flag in the segment makes the
 * others leave at their next MPI call; stragglers are killed after a grace period
 * (by PID: only the children this launcher started).  Exit status: 0 if every
 * rank returned 0, otherwise the first failing rank's status.
 */
#define _GNU_SOURCE
#include <errno.h>
#include <fcntl.h>
#include <signal.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/wait.h>
#include <time.h>
#include <unistd.h>
#include "minimpi_shm.h"

static char shm_name[64];
static pid_t *kids;
static int nkids;

static void cleanup(void)
{
   if (shm_name[0]) shm_unlink(shm_name);
}

static void on_signal(int sig)
{
   int i;
   for (i = 0; i < nkids; i++)
      if (kids[i] > 0) kill(kids[i], SIGTERM);
   cleanup();
   _exit(128 + sig);
}

int main(int argc, char **argv)
{
   int n = 0, a = 1, i, fd, status = 0, left;
   size_t ring = 0, total;
   minimpi_seg *seg;
   while (a < argc && argv[a][0] == '-') {
      if ((!strcmp(argv[a], "-n") || !strcmp(argv[a], "-np")) && a + 1 < argc) { n = atoi(argv[a + 1]); a += 2; }
      else if (!strcmp(argv[a], "--ring-kb") && a + 1 < argc) { ring = (size_t) atol(argv[a + 1])*1024; a += 2; }
      else break;
   }
   if (n <= 0 || a >= argc) {
      fprintf(stderr, "usage: minimpirun -n N [--ring-kb K] program [args...]\n");
      return 2;
   }
   if (!ring) {
      ring = 1 << 20;
      while (ring > 4096 && (size_t) n*n*ring > ((size_t) 512 << 20)) ring >>= 1;
   }
   { size_t p = 4096; while (p < ring) p <<= 1; ring = p; }
   total = minimpi_seg_bytes(n, ring);
   snprintf(shm_name, sizeof shm_name, "/minimpi_%ld_%ld", (long) getpid(), (long) time(NULL));
   fd = shm_open(shm_name, O_CREAT | O_EXCL | O_RDWR, 0600);
   if (fd < 0) { perror("minimpirun: shm_open"); return 2; }
   atexit(cleanup);
   if (ftruncate(fd, (off_t) total) != 0) { perror("minimpirun: ftruncate"); return 2; }
   seg = (minimpi_seg *) mmap(NULL, total, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
   if (seg == MAP_FAILED) { perror("minimpirun: mmap"); return 2; }
   close(fd);
   minimpi_seg_init(seg, n, ring);

   kids = (pid_t *) calloc((size_t) n, sizeof(pid_t));
   nkids = n;
   signal(SIGINT, on_signal);
   signal(SIGTERM, on_signal);
   for (i = 0; i < n; i++) {
      pid_t p = fork();
      if (p < 0) { perror("minimpirun: fork"); on_signal(SIGTERM); }
      if (p == 0) {
         char buf[32];
         setenv("MINIMPI_SHM", shm_name, 1);
         snprintf(buf, sizeof buf, "%d", i);
         setenv("MINIMPI_RANK", buf, 1);
         setenv("RANK", buf, 1);
         setenv("LOCAL_RANK", buf, 1);
         snprintf(buf, sizeof buf, "%d", n);
         setenv("MINIMPI_SIZE", buf, 1);
         setenv("WORLD_SIZE", buf, 1);
         shm_name[0] = 0;                      /* the child must not unlink */
         execvp(argv[a], argv + a);
         fprintf(stderr, "minimpirun: cannot execute %s: %s\n", argv[a], strerror(errno));
         _exit(127);
      }
      kids[i] = p;
   }
   left = n;
   while (left > 0) {
      int st;
      pid_t p = waitpid(-1, &st, 0);
      int code;
      if (p < 0) { if (errno == EINTR) continue; break; }
      for (i = 0; i < n && kids[i] != p; i++) ;
      if (i == n) continue;
      kids[i] = 0;
      left--;
      code = WIFEXITED(st) ? WEXITSTATUS(st) : 128 + WTERMSIG(st);
      if (code != 0) {
         if (!status) {
            status = code;
            fprintf(stderr, "minimpirun: rank %d ended with status %d; stopping the job\n", i, code);
         }
         atomic_store(&seg->abort_flag, 1);
         /* grace period, then stop whoever is still computing outside MPI */
         {
            int waited = 0;
            while (left > 0 && waited < 40) {
               pid_t q = waitpid(-1, &st, WNOHANG);
               if (q > 0) {
                  int k;
                  for (k = 0; k < n && kids[k] != q; k++) ;
                  if (k < n) { kids[k] = 0; left--; }
               } else {
                  struct timespec ts = {0, 50*1000*1000};
                  nanosleep(&ts, NULL);
                  waited++;
               }
            }
            for (i = 0; i < n; i++)
               if (kids[i] > 0) kill(kids[i], SIGKILL);
         }
      }
   }
   return status;
}
