// mem_write_bw.c — CPU-side write bandwidth into the memory of each NUMA node
// (tools/r02 evidence for "what bounds N concurrent drains on this host": the DMA
// writes of the checkpoint drain end up in the same DDR channels).
//   mem_write_bw [threads_per_node=16] [MiB_per_thread=512] [reps=3]
// For every (cpu node, memory node) pair: T threads pinned to the cpu node write
// buffers bound to the memory node, with regular stores (memset) and with
// non-temporal stores; prints one JSON line per pair.
// Build: gcc -O2 -mavx2 -pthread -o mem_write_bw mem_write_bw.c
#define _GNU_SOURCE
#include <immintrin.h>
#include <pthread.h>
#include <sched.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/syscall.h>
#include <time.h>
#include <unistd.h>

static double now(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return ts.tv_sec + ts.tv_nsec * 1e-9;
}

static int node_cpus(int node, int* cpus, int max) {
  char path[128], buf[4096];
  snprintf(path, sizeof(path), "/sys/devices/system/node/node%d/cpulist", node);
  FILE* f = fopen(path, "r");
  if (!f) return 0;
  if (!fgets(buf, sizeof(buf), f)) { fclose(f); return 0; }
  fclose(f);
  int n = 0;
  for (char* tok = strtok(buf, ",\n"); tok; tok = strtok(NULL, ",\n")) {
    int a, b;
    if (sscanf(tok, "%d-%d", &a, &b) == 2) { for (int c = a; c <= b && n < max; ++c) cpus[n++] = c; }
    else if (sscanf(tok, "%d", &a) == 1 && n < max) cpus[n++] = a;
  }
  return n;
}

typedef struct {
  int cpu, mem_node, nt_stores, reps;
  size_t bytes;
  pthread_barrier_t* bar;
  double t0, t1;
} job_t;

static void bind_mem(void* p, size_t n, int node) {
  unsigned long mask[16] = {0};
  mask[node / 64] |= 1ul << (node % 64);
  syscall(SYS_mbind, p, n, 2 /*MPOL_BIND*/, mask, 1024ul, 0u);
}

static void* worker(void* arg) {
  job_t* j = (job_t*)arg;
  cpu_set_t set;
  CPU_ZERO(&set);
  CPU_SET(j->cpu, &set);
  sched_setaffinity(0, sizeof(set), &set);
  uint8_t* buf = mmap(NULL, j->bytes, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
  if (buf == MAP_FAILED) { perror("mmap"); exit(1); }
  bind_mem(buf, j->bytes, j->mem_node);
  memset(buf, 1, j->bytes);  // fault in
  pthread_barrier_wait(j->bar);
  j->t0 = now();
  for (int r = 0; r < j->reps; ++r) {
    if (j->nt_stores) {
      __m256i v = _mm256_set1_epi8((char)(r + 2));
      for (size_t o = 0; o < j->bytes; o += 128) {
        _mm256_stream_si256((__m256i*)(buf + o), v);
        _mm256_stream_si256((__m256i*)(buf + o + 32), v);
        _mm256_stream_si256((__m256i*)(buf + o + 64), v);
        _mm256_stream_si256((__m256i*)(buf + o + 96), v);
      }
      _mm_sfence();
    } else {
      memset(buf, r + 2, j->bytes);
    }
  }
  j->t1 = now();
  pthread_barrier_wait(j->bar);
  munmap(buf, j->bytes);
  return NULL;
}

int main(int argc, char** argv) {
  int T = argc > 1 ? atoi(argv[1]) : 16;
  size_t mib = argc > 2 ? (size_t)atol(argv[2]) : 512;
  int reps = argc > 3 ? atoi(argv[3]) : 3;
  int nodes = 0;
  char path[128];
  for (; nodes < 16; ++nodes) {
    snprintf(path, sizeof(path), "/sys/devices/system/node/node%d", nodes);
    if (access(path, F_OK) != 0) break;
  }
  if (nodes == 0) nodes = 1;
  for (int cn = 0; cn < nodes; ++cn) {
    int cpus[1024];
    int nc = node_cpus(cn, cpus, 1024);
    if (nc == 0) { nc = (int)sysconf(_SC_NPROCESSORS_ONLN); for (int i = 0; i < nc; ++i) cpus[i] = i; }
    int t = T < nc ? T : nc;
    for (int mn = 0; mn < nodes; ++mn)
      for (int nt = 0; nt < 2; ++nt) {
        pthread_t th[1024];
        job_t jobs[1024];
        pthread_barrier_t bar;
        pthread_barrier_init(&bar, NULL, t);
        for (int i = 0; i < t; ++i) {
          jobs[i] = (job_t){cpus[i], mn, nt, reps, mib << 20, &bar, 0, 0};
          pthread_create(&th[i], NULL, worker, &jobs[i]);
        }
        double t0 = 1e300, t1 = 0;
        for (int i = 0; i < t; ++i) {
          pthread_join(th[i], NULL);
          if (jobs[i].t0 < t0) t0 = jobs[i].t0;
          if (jobs[i].t1 > t1) t1 = jobs[i].t1;
        }
        double gb = (double)t * (double)(mib << 20) * reps / 1e9;
        printf("{\"probe\": \"mem_write_bw\", \"cpu_node\": %d, \"mem_node\": %d, \"threads\": %d, "
               "\"stores\": \"%s\", \"GBps\": %.1f}\n",
               cn, mn, t, nt ? "non_temporal" : "memset", gb / (t1 - t0));
        fflush(stdout);
        pthread_barrier_destroy(&bar);
      }
  }
  return 0;
}
