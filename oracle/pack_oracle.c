/* ORACLE (test infrastructure, not product code).
 *
 * Plain-C restatement of the reference's copy loop
 *   dlrover/python/elastic_agent/torch/ckpt_saver.py:198-231
 * for byte ranges that are already host-resident: range i (nbytes[i] bytes at
 * src[i]) is copied to dst + off[i].  Used (a) by tests as an independent
 * check of the numpy oracle and (b) by bench.py's cpu_baseline leg as the
 * "what would the host cores do" figure, single- or multi-threaded.
 * Build: oracle/Makefile -> oracle/_ref/libpack_oracle.so
 */
#include <pthread.h>
#include <stdint.h>
#include <string.h>

typedef struct {
  uint8_t* dst;
  const uint8_t* const* src;
  const uint64_t* off;
  const uint64_t* nbytes;
  uint64_t first, last; /* range indices [first, last) */
} job_t;

static void* run(void* arg) {
  job_t* j = (job_t*)arg;
  for (uint64_t i = j->first; i < j->last; ++i)
    if (j->nbytes[i]) memcpy(j->dst + j->off[i], j->src[i], j->nbytes[i]);
  return 0;
}

/* returns 0, or -1 on bad arguments */
int oracle_pack(uint8_t* dst, uint64_t n, const uint8_t* const* src, const uint64_t* off,
                const uint64_t* nbytes, int threads) {
  if (!dst || (n && (!src || !off || !nbytes))) return -1;
  if (threads < 1) threads = 1;
  if (threads > 64) threads = 64;
  if ((uint64_t)threads > n) threads = n ? (int)n : 1;
  pthread_t th[64];
  job_t jobs[64];
  uint64_t per = (n + threads - 1) / threads;
  int started = 0;
  for (int t = 0; t < threads; ++t) {
    uint64_t a = (uint64_t)t * per, b = a + per > n ? n : a + per;
    if (a >= b) break;
    jobs[t] = (job_t){dst, src, off, nbytes, a, b};
    if (t == threads - 1 || pthread_create(&th[t], 0, run, &jobs[t]) != 0) {
      run(&jobs[t]); /* last slice (or failed spawn) on the calling thread */
      th[t] = 0;
    }
    started = t + 1;
  }
  for (int t = 0; t < started; ++t)
    if (th[t]) pthread_join(th[t], 0);
  return 0;
}

/* inverse: dst[i] <- src + off[i] */
int oracle_unpack(const uint8_t* src, uint64_t n, uint8_t* const* dst, const uint64_t* off,
                  const uint64_t* nbytes) {
  if (!src || (n && (!dst || !off || !nbytes))) return -1;
  for (uint64_t i = 0; i < n; ++i)
    if (nbytes[i]) memcpy(dst[i], src + off[i], nbytes[i]);
  return 0;
}
