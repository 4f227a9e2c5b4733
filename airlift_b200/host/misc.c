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

/* The reference sorts with klib's in-place MSD byte radix sort (insertion sort at <= 64 elements).  Where equal
 * keys can meet, the order they end up in is observable downstream, so the same permutation
 * is performed here: count, then cycle elements into their buckets starting from the lowest non-empty one. */
#define RS_SMALL 64

#define RADIX_IMPL(NAME, T, KEY) \
static void NAME##_ins(T *b, T *e) \
{ \
	T *i, *j; \
	for (i = b + 1; i < e; ++i) { \
		if (!(KEY(*i) < KEY(*(i - 1)))) continue; \
		T t = *i; \
		for (j = i; j > b && KEY(t) < KEY(*(j - 1)); --j) *j = *(j - 1); \
		*j = t; \
	} \
} \
static void NAME##_lvl(T *b, T *e, int sh) \
{ \
	size_t cnt[256]; T *hd[256], *tl[256], *i; int d; \
	memset(cnt, 0, sizeof(cnt)); \
	for (i = b; i != e; ++i) ++cnt[(KEY(*i) >> sh) & 0xff]; \
	for (d = 0, i = b; d < 256; ++d) hd[d] = i, i += cnt[d], tl[d] = i; \
	for (d = 0; d < 256;) { \
		int to; \
		if (hd[d] == tl[d]) { ++d; continue; } \
		to = (int)((KEY(*hd[d]) >> sh) & 0xff); \
		if (to == d) { ++hd[d]; continue; } \
		{ T carry = *hd[d], sw; \
		  do { sw = carry; carry = *hd[to]; *hd[to]++ = sw; to = (int)((KEY(carry) >> sh) & 0xff); } while (to != d); \
		  *hd[d]++ = carry; } \
	} \
	if (sh == 0) return; \
	sh = sh > 8 ? sh - 8 : 0; \
	for (d = 0, i = b; d < 256; i = tl[d], ++d) { \
		if (tl[d] - i > RS_SMALL) NAME##_lvl(i, tl[d], sh); \
		else if (tl[d] - i > 1) NAME##_ins(i, tl[d]); \
	} \
} \
void NAME(T *beg, T *end) \
{ \
	if (end - beg <= RS_SMALL) NAME##_ins(beg, end); \
	else NAME##_lvl(beg, end, 56); \
}

#define KEY_X(v) ((v).x)
#define KEY_ID(v) (v)
RADIX_IMPL(radix_sort_128x, mm128_t, KEY_X)
RADIX_IMPL(radix_sort_64, uint64_t, KEY_ID)

/* ---- per-fragment arena (see mm2b_priv.h) */

__thread mm_arena_t *mm_tls_arena = 0;

/* Standard-size chunks are recycled instead of going back to malloc (with short-lived worker threads glibc grows and
 * trims its per-thread heaps with mprotect() on every batch).  Each worker thread borrows a private pool for its
 * lifetime, so the hot path takes no lock; pools themselves live in a mutex-protected free list. */
#define ARENA_STD_CAP (16384 - sizeof(mm_arena_chunk_t) - 16)
#define ARENA_FIRST_CAP (8192 - sizeof(mm_arena_chunk_t) - 16) /* most short-read fragments fit their whole state in this */
#define POOL_MAX_CHUNKS (1 << 20)
struct mm_chunk_pool_s { struct mm_chunk_pool_s *next; mm_arena_chunk_t *chunks, *small; int n, n_small; };
__thread mm_chunk_pool_t *mm_tls_pool = 0;
static mm_chunk_pool_t *g_pools = 0;
static pthread_mutex_t g_pools_mu = PTHREAD_MUTEX_INITIALIZER;

mm_chunk_pool_t *mm_pool_acquire(void)
{
	mm_chunk_pool_t *p;
	pthread_mutex_lock(&g_pools_mu);
	if ((p = g_pools) != 0) g_pools = p->next;
	pthread_mutex_unlock(&g_pools_mu);
	if (p == 0) p = (mm_chunk_pool_t*)calloc(1, sizeof(*p));
	p->next = 0;
	return p;
}

void mm_pool_release(mm_chunk_pool_t *p)
{
	if (p == 0) return;
	pthread_mutex_lock(&g_pools_mu);
	p->next = g_pools, g_pools = p;
	pthread_mutex_unlock(&g_pools_mu);
}

static mm_arena_chunk_t *chunk_get(size_t cap)
{
	mm_arena_chunk_t *c = 0;
	mm_chunk_pool_t *p = mm_tls_pool;
	if (cap == ARENA_STD_CAP && p && p->chunks) c = p->chunks, p->chunks = c->next, --p->n;
	else if (cap == ARENA_FIRST_CAP && p && p->small) c = p->small, p->small = c->next, --p->n_small;
	if (c == 0) c = (mm_arena_chunk_t*)malloc(sizeof(mm_arena_chunk_t) + 16 + cap);
	c->cap = cap, c->used = 0, c->next = 0;
	return c;
}

void *mm_amalloc(size_t n)
{
	mm_arena_t *a = mm_tls_arena;
	mm_arena_chunk_t *c;
	if (a == 0) return malloc(n);
	n = (n + 15) & ~(size_t)15;
	c = a->head;
	if (c == 0 || c->used + n > c->cap) {
		mm_arena_chunk_t *nc = chunk_get(c == 0 && n <= ARENA_FIRST_CAP ? ARENA_FIRST_CAP : n > ARENA_STD_CAP ? n : ARENA_STD_CAP);
		if (c && n > ARENA_STD_CAP / 2) { nc->next = c->next; c->next = nc; c = nc; } /* a big block gets its own chunk; keep filling the current one */
		else { nc->next = c; a->head = nc; c = nc; }
	}
	{
		char *base = (char*)(((size_t)(c + 1) + 15) & ~(size_t)15);
		void *r = base + c->used;
		c->used += n;
		return r;
	}
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

void mm_arena_release(mm_arena_t *a)
{
	mm_arena_chunk_t *c = a->head, *n;
	mm_chunk_pool_t *p = mm_tls_pool;
	for (; c; c = n) {
		n = c->next;
		if (p && c->cap == ARENA_STD_CAP && p->n < POOL_MAX_CHUNKS) c->next = p->chunks, p->chunks = c, ++p->n;
		else if (p && c->cap == ARENA_FIRST_CAP && p->n_small < 4 * POOL_MAX_CHUNKS) c->next = p->small, p->small = c, ++p->n_small;
		else free(c);
	}
	a->head = 0;
}
