/* misc.c -- timers, checked output, lookup tables and the two order-defining sorts. */
#include <stdio.h>
#include <sys/resource.h>
#include <sys/time.h>
#include <pthread.h>
#include "mm2b_priv.h"

int mm_verbose = 1;
int mm_dbg_flag = 0;
double mm_realtime0;

double cputime(void)
{
	struct rusage r;
	getrusage(RUSAGE_SELF, &r);
	return r.ru_utime.tv_sec + r.ru_stime.tv_sec + 1e-6 * (r.ru_utime.tv_usec + r.ru_stime.tv_usec);
}

long peakrss(void)
{
	struct rusage r;
	getrusage(RUSAGE_SELF, &r);
	return r.ru_maxrss * 1024;
}

double realtime(void)
{
	struct timeval tp;
	gettimeofday(&tp, NULL);
	return tp.tv_sec + tp.tv_usec * 1e-6;
}

void mm_err_puts(const char *str)
{ /* one record per puts(); a short write is fatal (misc.c:123-131) */
	if (puts(str) == EOF) {
		perror("[ERROR] failed to write the results");
		exit(EXIT_FAILURE);
	}
}

/* base -> 0..3, anything else 4; U counts as T (sketch.c:9-26) */
unsigned char seq_nt4_table[256];
/* IUPAC complement, case preserved, other bytes unchanged (bseq.c:11-28) */
unsigned char seq_comp_table[256];

__attribute__((constructor)) static void init_tables(void)
{
	int i;
	const char *from = "ACGTUMRWSYKVHDBN", *to = "TGCAAKYWSRMBDHVN";
	for (i = 0; i < 256; ++i) seq_nt4_table[i] = i < 4 ? i : 4, seq_comp_table[i] = (unsigned char)i;
	seq_nt4_table['A'] = seq_nt4_table['a'] = 0; seq_nt4_table['C'] = seq_nt4_table['c'] = 1;
	seq_nt4_table['G'] = seq_nt4_table['g'] = 2; seq_nt4_table['T'] = seq_nt4_table['t'] = 3;
	seq_nt4_table['U'] = seq_nt4_table['u'] = 3;
	for (i = 0; from[i]; ++i) {
		seq_comp_table[(unsigned char)from[i]] = (unsigned char)to[i];
		seq_comp_table[(unsigned char)(from[i] + 32)] = (unsigned char)(to[i] + 32);
	}
}

RADIX_IMPL(radix_sort_128x, mm128_t, RS_KEY_X, RS_TABLES_STACK)
RADIX_IMPL(radix_sort_64, uint64_t, RS_KEY_ID, RS_TABLES_STACK)

/* ---- shard-lifetime bump arena (see mm2b_priv.h) */

__thread mm_arena_t *mm_tls_arena = 0;

/* Memory is carved from 256 KB blocks that are recycled through a global free list, so a mini-batch takes no trip to
 * malloc for its working state (short-lived worker threads otherwise make glibc grow and trim its per-thread heaps with
 * mprotect() all the time: that was 40 % of the host time).  A worker bumps inside its own block without locks; the block
 * it leaves half full when it exits is handed to the next worker of the same shard. */
#define ABLOCK_BYTES ((size_t)256 << 10)
#define ABLOCK_CAP (ABLOCK_BYTES - 32)          /* usable bytes after the 32-byte header slot */
#define ABLOCK_KEEP_MAX ((size_t)16 << 30)      /* blocks kept for reuse */
static mm_ablock_t *g_free_blocks = 0;
static size_t g_free_bytes = 0;
static pthread_mutex_t g_blocks_mu = PTHREAD_MUTEX_INITIALIZER;
static uint64_t g_arena_gen = 0;
static __thread struct { mm_arena_t *owner; uint64_t gen; char *cur, *end; } tls_bump;

void mm_arena_init(mm_arena_t *a)
{
	memset(a, 0, sizeof(*a));
	pthread_mutex_init(&a->mu, 0);
	a->gen = __sync_add_and_fetch(&g_arena_gen, 1);
}

static mm_ablock_t *block_new(mm_arena_t *a, size_t cap)
{ /* a standard block from the free list (or the system), or a dedicated one for a large request; registered with the arena */
	mm_ablock_t *b = 0;
	if (cap == ABLOCK_CAP) {
		pthread_mutex_lock(&g_blocks_mu);
		if ((b = g_free_blocks) != 0) g_free_blocks = b->next, g_free_bytes -= ABLOCK_BYTES;
		pthread_mutex_unlock(&g_blocks_mu);
	}
	if (b == 0) b = (mm_ablock_t*)malloc(32 + cap);
	b->cap = cap;
	pthread_mutex_lock(&a->mu);
	b->next = a->blocks, a->blocks = b;
	pthread_mutex_unlock(&a->mu);
	return b;
}

void *mm_amalloc(size_t n)
{
	mm_arena_t *a = mm_tls_arena;
	if (a == 0) return malloc(n);
	n = (n + 15) & ~(size_t)15;
	if (tls_bump.owner != a || tls_bump.gen != a->gen) tls_bump.owner = a, tls_bump.gen = a->gen, tls_bump.cur = tls_bump.end = 0;
	if (tls_bump.cur + n > tls_bump.end) {
		if (n > ABLOCK_CAP / 4) return (char*)block_new(a, n) + 32; /* own block; the current one keeps filling */
		pthread_mutex_lock(&a->mu); /* a block another worker left behind? */
		if (a->n_partial > 0 && a->partial[a->n_partial - 1].end - a->partial[a->n_partial - 1].cur >= (long)n) {
			--a->n_partial;
			tls_bump.cur = a->partial[a->n_partial].cur, tls_bump.end = a->partial[a->n_partial].end;
			pthread_mutex_unlock(&a->mu);
		} else {
			mm_ablock_t *b;
			pthread_mutex_unlock(&a->mu);
			b = block_new(a, ABLOCK_CAP);
			tls_bump.cur = (char*)b + 32, tls_bump.end = tls_bump.cur + ABLOCK_CAP;
		}
	}
	{
		void *r = tls_bump.cur;
		tls_bump.cur += n;
		return r;
	}
}

/* a worker thread is about to exit: pass on what is left of its block */
void mm_arena_thread_done(void)
{
	mm_arena_t *a = tls_bump.owner;
	if (a && tls_bump.end - tls_bump.cur >= 4096 && tls_bump.gen == a->gen) {
		pthread_mutex_lock(&a->mu);
		if (a->n_partial < MM_ARENA_MAX_PARTIAL) a->partial[a->n_partial].cur = tls_bump.cur, a->partial[a->n_partial].end = tls_bump.end, ++a->n_partial;
		pthread_mutex_unlock(&a->mu);
	}
	tls_bump.owner = 0, tls_bump.cur = tls_bump.end = 0;
}

void *mm_acalloc(size_t n, size_t sz)
{
	void *p;
	if (mm_tls_arena == 0) return calloc(n, sz);
	p = mm_amalloc(n * sz);
	memset(p, 0, n * sz);
	return p;
}

void *mm_arealloc(void *p, size_t old_bytes, size_t new_bytes)
{
	void *q;
	if (mm_tls_arena == 0) return realloc(p, new_bytes);
	q = mm_amalloc(new_bytes);
	if (p && old_bytes) memcpy(q, p, old_bytes < new_bytes ? old_bytes : new_bytes);
	return q;
}

void mm_afree(void *p) { if (mm_tls_arena == 0) free(p); }

/* the shard is done: every block goes back to the free list in one sweep */
void mm_arena_release(mm_arena_t *a)
{
	mm_ablock_t *b = a->blocks, *n, *keep = 0, *keep_tail = 0;
	size_t n_keep = 0;
	for (; b; b = n) {
		n = b->next;
		if (b->cap == ABLOCK_CAP) { b->next = keep; if (keep == 0) keep_tail = b; keep = b; ++n_keep; }
		else free(b);
	}
	if (keep) {
		pthread_mutex_lock(&g_blocks_mu);
		if (g_free_bytes + n_keep * ABLOCK_BYTES <= ABLOCK_KEEP_MAX) {
			keep_tail->next = g_free_blocks, g_free_blocks = keep, g_free_bytes += n_keep * ABLOCK_BYTES;
			keep = 0;
		}
		pthread_mutex_unlock(&g_blocks_mu);
		for (b = keep; b; b = n) { n = b->next; free(b); }
	}
	a->blocks = 0, a->n_partial = 0;
	a->gen = __sync_add_and_fetch(&g_arena_gen, 1);
	if (tls_bump.owner == a) tls_bump.owner = 0, tls_bump.cur = tls_bump.end = 0;
}

/* glibc: keep freed memory instead of trimming and re-growing the heaps on every mini-batch */
#include <malloc.h>
void mm_b200_tune_malloc(void)
{
	static int done = 0;
	if (__sync_lock_test_and_set(&done, 1)) return;
	mallopt(M_TRIM_THRESHOLD, 1 << 30);
	mallopt(M_TOP_PAD, 64 << 20);
	mallopt(M_MMAP_THRESHOLD, 32 << 20);
}
