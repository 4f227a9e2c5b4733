/* CPU ORACLE -- TEST INFRASTRUCTURE ONLY (see mm2_oracle.h).
 * Scalar C restatement of sketch / index lookup / seed collection / chaining /
 * ksw_extd2 of the AirLift minimap2 fork.  Paths cited as mm2/<file>:<line> are
 * /root/reference/src/minimap2-master_remapping/<file>. */
#include <stdlib.h>
#include <string.h>
#include <assert.h>
#include "mm2_oracle.h"

/* ---------------------------------------------------------------- sketch */

static const uint8_t nt4[256] = { /* mm2/sketch.c:9-26: A/a=0 C/c=1 G/g=2 T/t/U/u=3, else 4; bytes 0..3 map to themselves */
	0,1,2,3, 4,4,4,4, 4,4,4,4, 4,4,4,4,  4,4,4,4, 4,4,4,4, 4,4,4,4, 4,4,4,4,
	4,4,4,4, 4,4,4,4, 4,4,4,4, 4,4,4,4,  4,4,4,4, 4,4,4,4, 4,4,4,4, 4,4,4,4,
	4,0,4,1, 4,4,4,2, 4,4,4,4, 4,4,4,4,  4,4,4,4, 3,3,4,4, 4,4,4,4, 4,4,4,4,
	4,0,4,1, 4,4,4,2, 4,4,4,4, 4,4,4,4,  4,4,4,4, 3,3,4,4, 4,4,4,4, 4,4,4,4,
	4,4,4,4, 4,4,4,4, 4,4,4,4, 4,4,4,4,  4,4,4,4, 4,4,4,4, 4,4,4,4, 4,4,4,4,
	4,4,4,4, 4,4,4,4, 4,4,4,4, 4,4,4,4,  4,4,4,4, 4,4,4,4, 4,4,4,4, 4,4,4,4,
	4,4,4,4, 4,4,4,4, 4,4,4,4, 4,4,4,4,  4,4,4,4, 4,4,4,4, 4,4,4,4, 4,4,4,4,
	4,4,4,4, 4,4,4,4, 4,4,4,4, 4,4,4,4,  4,4,4,4, 4,4,4,4, 4,4,4,4, 4,4,4,4
};

static inline uint64_t inv_hash(uint64_t key, uint64_t mask) /* mm2/sketch.c:28-38 */
{
	key = (~key + (key << 21)) & mask;
	key ^= key >> 24;
	key = (key + (key << 3) + (key << 8)) & mask;
	key ^= key >> 14;
	key = (key + (key << 2) + (key << 4)) & mask;
	key ^= key >> 28;
	key = (key + (key << 31)) & mask;
	return key;
}

#define ORC_NONE UINT64_MAX

int orc_sketch(const char *str, int len, int w, int k, uint32_t rid, int is_hpc, orc128_t *out, int cap)
{ /* mm2/sketch.c:77-143, one base per iteration */
	const uint64_t shift1 = 2 * (k - 1), mask = (1ULL << 2 * k) - 1;
	uint64_t kmer[2] = {0, 0};
	orc128_t ring[256], best = {ORC_NONE, ORC_NONE};
	int hp_q[32], hp_front = 0, hp_count = 0; /* run lengths of the last k homopolymers (tiny_queue_t, sketch.c:40-58) */
	int i, j, l = 0, ring_pos = 0, best_pos = 0, span = 0, n = 0;
#define EMIT(v) do { if (n < cap) out[n] = (v); ++n; } while (0)
	assert(len > 0 && w > 0 && w < 256 && k > 0 && k <= 28);
	memset(ring, 0xff, sizeof(orc128_t) * w);
	for (i = 0; i < len; ++i) {
		int c = nt4[(uint8_t)str[i]];
		orc128_t info = {ORC_NONE, ORC_NONE};
		if (c < 4) {
			int z;
			if (is_hpc) { /* sketch.c:94-104 */
				int run = 1;
				if (i + 1 < len && nt4[(uint8_t)str[i + 1]] == c) {
					for (run = 2; i + run < len; ++run)
						if (nt4[(uint8_t)str[i + run]] != c) break;
					i += run - 1;
				}
				hp_q[(hp_count++ + hp_front) & 0x1f] = run;
				span += run;
				if (hp_count > k) { span -= hp_q[hp_front++]; hp_front &= 0x1f; --hp_count; }
			} else span = l + 1 < k ? l + 1 : k;
			kmer[0] = (kmer[0] << 2 | c) & mask;
			kmer[1] = (kmer[1] >> 2) | (3ULL ^ c) << shift1;
			if (kmer[0] == kmer[1]) continue; /* palindromic k-mer: strand unknown (sketch.c:108) */
			z = kmer[0] < kmer[1] ? 0 : 1;
			++l;
			if (l >= k && span < 256) {
				info.x = inv_hash(kmer[z], mask) << 8 | span;
				info.y = (uint64_t)rid << 32 | (uint32_t)i << 1 | z;
			}
		} else l = 0, hp_count = hp_front = 0, span = 0;
		ring[ring_pos] = info;
		if (l == w + k - 1 && best.x != ORC_NONE) { /* first full window: flush k-mers identical to the min (sketch.c:117-122) */
			for (j = ring_pos + 1; j < w; ++j) if (best.x == ring[j].x && ring[j].y != best.y) EMIT(ring[j]);
			for (j = 0; j < ring_pos; ++j)     if (best.x == ring[j].x && ring[j].y != best.y) EMIT(ring[j]);
		}
		if (info.x <= best.x) { /* new minimum (sketch.c:123-125) */
			if (l >= w + k && best.x != ORC_NONE) EMIT(best);
			best = info, best_pos = ring_pos;
		} else if (ring_pos == best_pos) { /* the minimum left the window (sketch.c:126-137) */
			if (l >= w + k - 1 && best.x != ORC_NONE) EMIT(best);
			for (j = ring_pos + 1, best.x = ORC_NONE; j < w; ++j) if (best.x >= ring[j].x) best = ring[j], best_pos = j;
			for (j = 0; j <= ring_pos; ++j)                       if (best.x >= ring[j].x) best = ring[j], best_pos = j;
			if (l >= w + k - 1 && best.x != ORC_NONE) {
				for (j = ring_pos + 1; j < w; ++j) if (best.x == ring[j].x && best.y != ring[j].y) EMIT(ring[j]);
				for (j = 0; j <= ring_pos; ++j)    if (best.x == ring[j].x && best.y != ring[j].y) EMIT(ring[j]);
			}
		}
		if (++ring_pos == w) ring_pos = 0;
	}
	if (best.x != ORC_NONE) EMIT(best);
#undef EMIT
	return n;
}

/* ------------------------------------------------------------ klib sorts */
/* In-place MSD byte radix sort with insertion sort at <=64 elements.  Tie order
 * among equal keys is part of the observable behaviour, so
 * the permutation cycle of mm2/ksort.h:116-146 is followed step for step. */
#define RS_SMALL 64

#define DEFINE_RADIX(NAME, T, KEY)                                                            \
static void NAME##_insertion(T *beg, T *end)                                                  \
{                                                                                             \
	T *i;                                                                                     \
	for (i = beg + 1; i < end; ++i)                                                           \
		if (KEY(*i) < KEY(*(i - 1))) {                                                        \
			T *j, tmp = *i;                                                                   \
			for (j = i; j > beg && KEY(tmp) < KEY(*(j - 1)); --j) *j = *(j - 1);              \
			*j = tmp;                                                                         \
		}                                                                                     \
}                                                                                             \
static void NAME##_msd(T *beg, T *end, int shift)                                             \
{                                                                                             \
	T *head[256], *tail[256], *i;                                                             \
	int d;                                                                                    \
	size_t cnt[256];                                                                          \
	memset(cnt, 0, sizeof(cnt));                                                              \
	for (i = beg; i != end; ++i) ++cnt[KEY(*i) >> shift & 0xff];                              \
	head[0] = beg, tail[0] = beg + cnt[0];                                                    \
	for (d = 1; d < 256; ++d) head[d] = tail[d - 1], tail[d] = head[d] + cnt[d];              \
	for (d = 0; d < 256;) { /* cycle-leader permutation (ksort.h:126-138) */                  \
		if (head[d] != tail[d]) {                                                             \
			int to = (int)(KEY(*head[d]) >> shift & 0xff);                                    \
			if (to != d) {                                                                    \
				T carry = *head[d], swap;                                                     \
				do {                                                                          \
					swap = carry; carry = *head[to]; *head[to]++ = swap;                      \
					to = (int)(KEY(carry) >> shift & 0xff);                                   \
				} while (to != d);                                                            \
				*head[d]++ = carry;                                                           \
			} else ++head[d];                                                                 \
		} else ++d;                                                                           \
	}                                                                                         \
	if (shift) {                                                                              \
		T *lo = beg;                                                                          \
		shift = shift > 8 ? shift - 8 : 0;                                                    \
		for (d = 0; d < 256; ++d) {                                                           \
			if (tail[d] - lo > RS_SMALL) NAME##_msd(lo, tail[d], shift);                      \
			else if (tail[d] - lo > 1) NAME##_insertion(lo, tail[d]);                         \
			lo = tail[d];                                                                     \
		}                                                                                     \
	}                                                                                         \
}                                                                                             \
void NAME(T *beg, T *end)                                                                     \
{                                                                                             \
	if (end - beg <= RS_SMALL) NAME##_insertion(beg, end);                                    \
	else NAME##_msd(beg, end, 56);                                                            \
}

#define KEY128(a) ((a).x)
#define KEY64(a) (a)
DEFINE_RADIX(orc_radix_sort_128x, orc128_t, KEY128)
DEFINE_RADIX(orc_radix_sort_64, uint64_t, KEY64)

/* ------------------------------------------------------------------ index */

struct orc_idx_s {
	int w, k, is_hpc;
	int64_t n;        /* minimizer occurrences */
	uint64_t *key;    /* hash value (x>>8), ascending */
	uint64_t *pos;    /* rid<<32 | lastPos<<1 | strand, ascending within a key (index.c:230) */
	int64_t n_keys;
	int64_t *key_off; /* first occurrence of each distinct key; key_off[n_keys] = n */
};

static int cmp_kp(const void *a, const void *b)
{
	const orc128_t *p = (const orc128_t*)a, *q = (const orc128_t*)b;
	if (p->x != q->x) return p->x < q->x ? -1 : 1;
	return p->y < q->y ? -1 : p->y > q->y;
}

orc_idx_t *orc_idx_build(int w, int k, int is_hpc, int n_seq, const char **seqs)
{ /* index.c:385-432 (sketch every sequence) + index.c:191-243 (group by minimizer, positions ascending) */
	orc_idx_t *mi = (orc_idx_t*)calloc(1, sizeof(*mi));
	orc128_t *all = 0;
	int64_t n = 0, m = 0, i;
	mi->w = w, mi->k = k, mi->is_hpc = is_hpc;
	for (i = 0; i < n_seq; ++i) {
		int len = (int)strlen(seqs[i]), got;
		if (len == 0) continue;
		if (n + len > m) { m = (n + len) * 2; all = (orc128_t*)realloc(all, m * sizeof(orc128_t)); }
		got = orc_sketch(seqs[i], len, w, k, (uint32_t)i, is_hpc, all + n, len);
		n += got;
	}
	for (i = 0; i < n; ++i) all[i].x >>= 8;
	qsort(all, n, sizeof(orc128_t), cmp_kp);
	mi->n = n;
	mi->key = (uint64_t*)malloc((n + 1) * 8);
	mi->pos = (uint64_t*)malloc((n + 1) * 8);
	mi->key_off = (int64_t*)malloc((n + 2) * 8);
	for (i = 0; i < n; ++i) {
		mi->key[i] = all[i].x, mi->pos[i] = all[i].y;
		if (i == 0 || all[i].x != all[i - 1].x) mi->key_off[mi->n_keys++] = i;
	}
	mi->key_off[mi->n_keys] = n;
	free(all);
	return mi;
}

void orc_idx_destroy(orc_idx_t *mi)
{
	if (!mi) return;
	free(mi->key); free(mi->pos); free(mi->key_off); free(mi);
}

int64_t orc_idx_n_minimizers(const orc_idx_t *mi) { return mi->n; }

const uint64_t *orc_idx_get(const orc_idx_t *mi, uint64_t minier, int *n)
{ /* index.c:81-98: NULL/0 when absent, else the ascending position list */
	int64_t lo = 0, hi = mi->n_keys - 1;
	*n = 0;
	while (lo <= hi) {
		int64_t mid = (lo + hi) >> 1;
		uint64_t kx = mi->key[mi->key_off[mid]];
		if (kx < minier) lo = mid + 1;
		else if (kx > minier) hi = mid - 1;
		else {
			*n = (int)(mi->key_off[mid + 1] - mi->key_off[mid]);
			return &mi->pos[mi->key_off[mid]];
		}
	}
	return 0;
}

static int cmp_u32(const void *a, const void *b)
{
	uint32_t x = *(const uint32_t*)a, y = *(const uint32_t*)b;
	return x < y ? -1 : x > y;
}

int32_t orc_idx_cal_max_occ(const orc_idx_t *mi, float f)
{ /* index.c:164-185: (1-f) quantile of per-minimizer occurrence counts, plus one */
	uint32_t *cnt, thres;
	int64_t i;
	if (f <= 0.) return INT32_MAX;
	if (mi->n_keys == 0) return 1;
	cnt = (uint32_t*)malloc(mi->n_keys * 4);
	for (i = 0; i < mi->n_keys; ++i) cnt[i] = (uint32_t)(mi->key_off[i + 1] - mi->key_off[i]);
	qsort(cnt, mi->n_keys, 4, cmp_u32);
	thres = cnt[(uint32_t)((1. - f) * mi->n_keys)] + 1;
	free(cnt);
	return (int32_t)thres;
}

/* ----------------------------------------------------------- seed collection */

#define F_FOR_ONLY 0x100000
#define F_REV_ONLY 0x200000
#define SEED_TANDEM (1ULL << 42)
#define SEED_SEG_SHIFT 48
#define SEED_SEG_MASK (0xffULL << SEED_SEG_SHIFT)

typedef struct { uint32_t n, q_pos, q_span, seg_id, is_tandem; const uint64_t *cr; } match_t;

static int skip_strand(int64_t flag, uint64_t r, const match_t *q)
{ /* map.c:139-145 */
	if (flag & (F_FOR_ONLY | F_REV_ONLY)) {
		if ((r & 1) == (q->q_pos & 1)) { if (flag & F_REV_ONLY) return 1; }
		else if (flag & F_FOR_ONLY) return 1;
	}
	return 0;
}

static inline void make_anchor(orc128_t *p, uint64_t r, const match_t *q, int qlen)
{ /* map.c:176-187 == map.c:232-241 */
	int32_t rpos = (uint32_t)r >> 1;
	if ((r & 1) == (q->q_pos & 1)) {
		p->x = (r & 0xffffffff00000000ULL) | rpos;
		p->y = (uint64_t)q->q_span << 32 | q->q_pos >> 1;
	} else {
		p->x = 1ULL << 63 | (r & 0xffffffff00000000ULL) | rpos;
		p->y = (uint64_t)q->q_span << 32 | (uint32_t)(qlen - ((q->q_pos >> 1) + 1 - q->q_span) - 1);
	}
	p->y |= (uint64_t)q->seg_id << SEED_SEG_SHIFT;
	if (q->is_tandem) p->y |= SEED_TANDEM;
}

/* binary max... min-heap on x with klib's exact sift-down (ksort.h:43-53, heap_lt = a.x > b.x) */
static void heap_down(size_t i, size_t n, orc128_t *l)
{
	size_t k = i;
	orc128_t tmp = l[i];
	while ((k = (k << 1) + 1) < n) {
		if (k != n - 1 && l[k].x > l[k + 1].x) ++k;
		if (l[k].x > tmp.x) break;
		l[i] = l[k]; i = k;
	}
	l[i] = tmp;
}

int64_t orc_collect_seeds(const orc_idx_t *mi, int heap_sort, int64_t flag, int max_occ, int n_mv, const orc128_t *mv,
                          int qlen, orc128_t *a, int *rep_len, int *n_mini_pos, uint64_t *mini_pos)
{
	match_t *m = (match_t*)malloc((n_mv + 1) * sizeof(match_t));
	int i, n_m = 0, rep_st = 0, rep_en = 0, n_mp = 0;
	int64_t n_a = 0;
	*rep_len = 0;
	for (i = 0; i < n_mv; ++i) { /* collect_matches, map.c:90-123 */
		const orc128_t *p = &mv[i];
		uint32_t q_pos = (uint32_t)p->y, q_span = p->x & 0xff;
		int t;
		const uint64_t *cr = orc_idx_get(mi, p->x >> 8, &t);
		if (t >= max_occ) {
			int en = (q_pos >> 1) + 1, st = en - q_span;
			if (st > rep_en) { *rep_len += rep_en - rep_st; rep_st = st, rep_en = en; }
			else rep_en = en;
		} else {
			match_t *q = &m[n_m++];
			q->q_pos = q_pos, q->q_span = q_span, q->cr = cr, q->n = t, q->seg_id = (uint32_t)(p->y >> 32);
			q->is_tandem = 0;
			if (i > 0 && p->x >> 8 == mv[i - 1].x >> 8) q->is_tandem = 1;
			if (i < n_mv - 1 && p->x >> 8 == mv[i + 1].x >> 8) q->is_tandem = 1;
			n_a += q->n;
			if (mini_pos) mini_pos[n_mp] = (uint64_t)q_span << 32 | q_pos >> 1;
			++n_mp;
		}
	}
	*rep_len += rep_en - rep_st;
	if (n_mini_pos) *n_mini_pos = n_mp;
	if (a == 0) { free(m); return n_a; }

	if (heap_sort) { /* collect_seed_hits_heap, map.c:149-213 */
		orc128_t *heap = (orc128_t*)malloc((n_m + 1) * sizeof(orc128_t));
		size_t hs = 0;
		int64_t n_for = 0, n_rev = 0, j;
		for (i = 0; i < n_m; ++i)
			if (m[i].n > 0) { heap[hs].x = m[i].cr[0]; heap[hs].y = (uint64_t)i << 32; ++hs; }
		if (hs > 1) for (j = (int64_t)(hs >> 1) - 1; j >= 0; --j) heap_down((size_t)j, hs, heap); /* ks_heapmake */
		while (hs > 0) {
			match_t *q = &m[heap->y >> 32];
			uint64_t r = heap->x;
			if (!skip_strand(flag, r, q)) {
				if ((r & 1) == (q->q_pos & 1)) make_anchor(&a[n_for++], r, q, qlen);
				else make_anchor(&a[n_a - (++n_rev)], r, q, qlen);
			}
			if ((uint32_t)heap->y < q->n - 1) {
				++heap[0].y;
				heap[0].x = m[heap[0].y >> 32].cr[(uint32_t)heap[0].y];
			} else {
				heap[0] = heap[hs - 1];
				--hs;
			}
			heap_down(0, hs, heap);
		}
		free(heap);
		for (j = 0; j < n_rev >> 1; ++j) { /* reverse-strand anchors were written back to front */
			orc128_t t = a[n_a - 1 - j];
			a[n_a - 1 - j] = a[n_a - (n_rev - j)];
			a[n_a - (n_rev - j)] = t;
		}
		if (n_a > n_for + n_rev) {
			memmove(a + n_for, a + n_a - n_rev, n_rev * sizeof(orc128_t));
			n_a = n_for + n_rev;
		}
	} else { /* collect_seed_hits, map.c:215-247 */
		int64_t o = 0;
		for (i = 0; i < n_m; ++i) {
			uint32_t k;
			for (k = 0; k < m[i].n; ++k) {
				if (skip_strand(flag, m[i].cr[k], &m[i])) continue;
				make_anchor(&a[o++], m[i].cr[k], &m[i], qlen);
			}
		}
		n_a = o;
		orc_radix_sort_128x(a, a + n_a);
	}
	free(m);
	return n_a;
}

/* ---------------------------------------------------------------- chaining */

static inline int ilog2_u32(uint32_t v) /* chain.c:15-20; -1 for 0 is never used (guarded at chain.c:64) */
{
	int r = -1;
	while (v) ++r, v >>= 1;
	return r;
}

int orc_chain_dp(int max_dist_x, int max_dist_y, int bw, int max_skip, int max_iter, int min_cnt, int min_sc,
                 int is_cdna, int n_segs, int64_t n, orc128_t *a, uint64_t *u_out)
{
	int32_t *f, *p, *t, *v, n_u, n_v, k;
	int64_t i, j, st = 0;
	uint64_t *u, *u2, sum_qspan = 0;
	float avg_qspan;
	orc128_t *b, *w;
	if (n == 0 || a == 0) return 0;
	f = (int32_t*)malloc(n * 4); p = (int32_t*)malloc(n * 4);
	t = (int32_t*)calloc(n, 4);  v = (int32_t*)malloc(n * 4);
	for (i = 0; i < n; ++i) sum_qspan += a[i].y >> 32 & 0xff;
	avg_qspan = (float)sum_qspan / n; /* chain.c:42 */

	for (i = 0; i < n; ++i) { /* chain.c:45-85 */
		uint64_t ri = a[i].x;
		int64_t max_j = -1;
		int32_t qi = (int32_t)a[i].y, q_span = a[i].y >> 32 & 0xff;
		int32_t max_f = q_span, n_skip = 0, min_d;
		int32_t sidi = (a[i].y & SEED_SEG_MASK) >> SEED_SEG_SHIFT;
		while (st < i && ri > a[st].x + max_dist_x) ++st;
		if (i - st > max_iter) st = i - max_iter;
		for (j = i - 1; j >= st; --j) {
			int64_t dr = ri - a[j].x;
			int32_t dq = qi - (int32_t)a[j].y, dd, sc, log_dd;
			int32_t sidj = (a[j].y & SEED_SEG_MASK) >> SEED_SEG_SHIFT;
			if ((sidi == sidj && dr == 0) || dq <= 0) continue;
			if ((sidi == sidj && dq > max_dist_y) || dq > max_dist_x) continue;
			dd = dr > dq ? dr - dq : dq - dr;
			if (sidi == sidj && dd > bw) continue;
			if (n_segs > 1 && !is_cdna && sidi == sidj && dr > max_dist_y) continue;
			min_d = dq < dr ? dq : dr;
			sc = min_d > q_span ? q_span : dq < dr ? dq : dr;
			log_dd = dd ? ilog2_u32(dd) : 0;
			if (is_cdna || sidi != sidj) {
				int c_log, c_lin;
				c_lin = (int)(dd * .01 * avg_qspan);
				c_log = log_dd;
				if (sidi != sidj && dr == 0) ++sc;
				else if (dr > dq || sidi != sidj) sc -= c_lin < c_log ? c_lin : c_log;
				else sc -= c_lin + (c_log >> 1);
			} else sc -= (int)(dd * .01 * avg_qspan) + (log_dd >> 1);
			sc += f[j];
			if (sc > max_f) {
				max_f = sc, max_j = j;
				if (n_skip > 0) --n_skip;
			} else if (t[j] == i) {
				if (++n_skip > max_skip) break;
			}
			if (p[j] >= 0) t[p[j]] = i;
		}
		f[i] = max_f, p[i] = max_j;
		v[i] = max_j >= 0 && v[max_j] > max_f ? v[max_j] : max_f;
	}

	memset(t, 0, n * 4); /* chain.c:88-111: chain ends, peaks, sort by peak score */
	for (i = 0; i < n; ++i) if (p[i] >= 0) t[p[i]] = 1;
	for (i = n_u = 0; i < n; ++i) if (t[i] == 0 && v[i] >= min_sc) ++n_u;
	if (n_u == 0) { free(f); free(p); free(t); free(v); return 0; }
	u = (uint64_t*)malloc(n_u * 8);
	for (i = n_u = 0; i < n; ++i)
		if (t[i] == 0 && v[i] >= min_sc) {
			j = i;
			while (j >= 0 && f[j] < v[j]) j = p[j];
			if (j < 0) j = i;
			u[n_u++] = (uint64_t)f[j] << 32 | j;
		}
	orc_radix_sort_64(u, u + n_u);
	for (i = 0; i < n_u >> 1; ++i) { uint64_t x = u[i]; u[i] = u[n_u - i - 1], u[n_u - i - 1] = x; }

	memset(t, 0, n * 4); /* chain.c:114-130: backtrack, stop at anchors already used */
	for (i = n_v = k = 0; i < n_u; ++i) {
		int32_t n_v0 = n_v, k0 = k;
		j = (int32_t)u[i];
		do { v[n_v++] = j; t[j] = 1; j = p[j]; } while (j >= 0 && t[j] == 0);
		if (j < 0) {
			if (n_v - n_v0 >= min_cnt) u[k++] = u[i] >> 32 << 32 | (n_v - n_v0);
		} else if ((int32_t)(u[i] >> 32) - f[j] >= min_sc) {
			if (n_v - n_v0 >= min_cnt) u[k++] = ((u[i] >> 32) - f[j]) << 32 | (n_v - n_v0);
		}
		if (k0 == k) n_v = n_v0;
	}
	n_u = k;

	b = (orc128_t*)malloc((n_v + 1) * sizeof(orc128_t)); /* chain.c:136-141 */
	for (i = 0, k = 0; i < n_u; ++i) {
		int32_t k0 = k, ni = (int32_t)u[i];
		for (j = 0; j < ni; ++j) b[k] = a[v[k0 + (ni - j - 1)]], ++k;
	}
	w = (orc128_t*)malloc((n_u + 1) * sizeof(orc128_t)); /* chain.c:144-160: order chains by first-anchor x */
	for (i = k = 0; i < n_u; ++i) {
		w[i].x = b[k].x, w[i].y = (uint64_t)k << 32 | i;
		k += (int32_t)u[i];
	}
	orc_radix_sort_128x(w, w + n_u);
	u2 = (uint64_t*)malloc((n_u + 1) * 8);
	for (i = k = 0; i < n_u; ++i) {
		int32_t jj = (int32_t)w[i].y, nn = (int32_t)u[jj];
		u2[i] = u[jj];
		memcpy(&a[k], &b[w[i].y >> 32], nn * sizeof(orc128_t));
		k += nn;
	}
	memcpy(u_out, u2, n_u * 8);
	free(f); free(p); free(t); free(v); free(u); free(u2); free(b); free(w);
	return n_u;
}

/* ---------------------------------------------------------------- ksw_extd2 */

#define NEG_INF (-0x40000000)
#define EZ_SCORE_ONLY 0x01
#define EZ_RIGHT      0x02
#define EZ_GENERIC_SC 0x04
#define EZ_APPROX_MAX 0x08
#define EZ_APPROX_DROP 0x10
#define EZ_EXTZ_ONLY  0x40
#define EZ_REV_CIGAR  0x80

static void push_cigar(orc_extz_t *ez, uint32_t op, int len) /* ksw2.h:103-113 */
{
	if (ez->n_cigar == 0 || op != (ez->cigar[ez->n_cigar - 1] & 0xf)) {
		if (ez->n_cigar == ez->m_cigar) {
			ez->m_cigar = ez->m_cigar ? ez->m_cigar << 1 : 4;
			ez->cigar = (uint32_t*)realloc(ez->cigar, (size_t)ez->m_cigar << 2);
		}
		ez->cigar[ez->n_cigar++] = len << 4 | op;
	} else ez->cigar[ez->n_cigar - 1] += len << 4;
}

static void backtrack(orc_extz_t *ez, int is_rev, const uint8_t *p, const int *off, const int *off_end, int n_col, int i0, int j0)
{ /* ksw2.h:119-151 with is_rot=1, min_intron_len=0 */
	int i = i0, j = j0, r, state = 0, k;
	ez->n_cigar = 0;
	while (i >= 0 && j >= 0) {
		int force_state = -1;
		uint32_t tmp;
		r = i + j;
		if (i < off[r]) force_state = 2;
		if (i > off_end[r]) force_state = 1;
		tmp = force_state < 0 ? p[(size_t)r * n_col + i - off[r]] : 0;
		if (state == 0) state = tmp & 7;
		else if (!(tmp >> (state + 2) & 1)) state = 0;
		if (state == 0) state = tmp & 7;
		if (force_state >= 0) state = force_state;
		if (state == 0) push_cigar(ez, 0, 1), --i, --j;
		else if (state == 1 || state == 3) push_cigar(ez, 2, 1), --i;
		else push_cigar(ez, 1, 1), --j;
	}
	if (i >= 0) push_cigar(ez, 2, i + 1);
	if (j >= 0) push_cigar(ez, 1, j + 1);
	if (!is_rev)
		for (k = 0; k < ez->n_cigar >> 1; ++k) {
			uint32_t t = ez->cigar[k]; ez->cigar[k] = ez->cigar[ez->n_cigar - 1 - k]; ez->cigar[ez->n_cigar - 1 - k] = t;
		}
}

static int apply_zdrop(orc_extz_t *ez, int32_t H, int r, int t, int zdrop, int8_t e) /* ksw2.h:160-176, is_rot=1 */
{
	if (H > (int32_t)ez->max) {
		ez->max = H, ez->max_t = t, ez->max_q = r - t;
	} else if (t >= ez->max_t && r - t >= ez->max_q) {
		int tl = t - ez->max_t, ql = (r - t) - ez->max_q, l;
		l = tl > ql ? tl - ql : ql - tl;
		if (zdrop >= 0 && (int32_t)ez->max - H > zdrop + l * e) { ez->zdropped = 1; return 1; }
	}
	return 0;
}

int64_t orc_ksw_band_cells(int qlen, int tlen, int w)
{
	int64_t cells = 0;
	int r;
	if (w < 0) w = tlen > qlen ? tlen : qlen;
	for (r = 0; r < qlen + tlen - 1; ++r) {
		int st = 0, en = tlen - 1;
		if (st < r - qlen + 1) st = r - qlen + 1;
		if (en > r) en = r;
		if (st < (r - w + 1) >> 1) st = (r - w + 1) >> 1;
		if (en > (r + w) >> 1) en = (r + w) >> 1;
		if (st > en) break;
		cells += en - st + 1;
	}
	return cells;
}

void orc_ksw_extd2(int qlen, const uint8_t *query, int tlen, const uint8_t *target, int8_t m, const int8_t *mat,
                   int8_t q, int8_t e, int8_t q2, int8_t e2, int w, int zdrop, int end_bonus, int flag, orc_extz_t *ez)
{
	/* One int8 lane per byte, 16-lane blocks processed exactly as the SSE4.1 code does
	 * (ksw2_extd2_sse.c:123-378): widened [st,en], stale lanes, over-reads and all.
	 * The seven lane arrays, the target copy sf[] and the reversed query qr[] live in ONE
	 * zero-filled buffer in the reference's order (ksw2_extd2_sse.c:99-102), so that reads
	 * running past sf[] see qr[] and reads past qr[] see zeros, as in the reference. */
	int r, t, qe, qe_pre = q + e /* taken before the penalty swap, ksw2_extd2_sse.c:60 vs :70 */, n_col_, tlen_, qlen_, last_st, last_en, max_sc, min_sc, long_thres, long_diff;
	int with_cigar = !(flag & EZ_SCORE_ONLY), approx_max = !!(flag & EZ_APPROX_MAX);
	int32_t *H = 0, H0 = 0, last_H0_t = 0;
	int8_t *mem, *u, *v, *x, *y, *x2, *y2, *s;
	uint8_t *sf, *qr, *pm = 0;
	int *off = 0, *off_end = 0;
	int8_t sc_mch, sc_mis, sc_N;
	size_t mem_sz;

	ez->max_q = ez->max_t = ez->mqe_t = ez->mte_q = -1; /* ksw_reset_extz, ksw2.h:153-158 */
	ez->max = 0, ez->score = ez->mqe = ez->mte = NEG_INF;
	ez->n_cigar = 0, ez->zdropped = 0, ez->reach_end = 0;
	if (m <= 1 || qlen <= 0 || tlen <= 0) return;
	if (q2 + e2 < q + e) { int8_t x_; x_ = q, q = q2, q2 = x_; x_ = e, e = e2, e2 = x_; }
	qe = q + e;
	sc_mch = mat[0], sc_mis = mat[1];
	sc_N = mat[m * m - 1] == 0 ? -e2 : mat[m * m - 1];
	if (w < 0) w = tlen > qlen ? tlen : qlen;
	tlen_ = (tlen + 15) / 16;
	n_col_ = qlen < tlen ? qlen : tlen;
	n_col_ = ((n_col_ < w + 1 ? n_col_ : w + 1) + 15) / 16 + 1;
	qlen_ = (qlen + 15) / 16;
	for (t = 1, max_sc = mat[0], min_sc = mat[1]; t < m * m; ++t) {
		max_sc = max_sc > mat[t] ? max_sc : mat[t];
		min_sc = min_sc < mat[t] ? min_sc : mat[t];
	}
	if (-min_sc > 2 * (q + e)) return;
	long_thres = e != e2 ? (q2 - q) / (e - e2) - 1 : 0;
	if (q2 + e2 + long_thres * e2 > q + e + long_thres * e) ++long_thres;
	long_diff = long_thres * (e - e2) - (q2 - q) - e2;

	mem_sz = ((size_t)tlen_ * 8 + qlen_ + 1) * 16;
	mem = (int8_t*)calloc(mem_sz + 64, 1);
	u = mem, v = u + tlen_ * 16, x = v + tlen_ * 16, y = x + tlen_ * 16, x2 = y + tlen_ * 16, y2 = x2 + tlen_ * 16;
	s = y2 + tlen_ * 16, sf = (uint8_t*)(s + tlen_ * 16), qr = sf + tlen_ * 16;
	memset(u, -q - e, tlen_ * 16); memset(v, -q - e, tlen_ * 16);
	memset(x, -q - e, tlen_ * 16); memset(y, -q - e, tlen_ * 16);
	memset(x2, -q2 - e2, tlen_ * 16); memset(y2, -q2 - e2, tlen_ * 16);
	if (!approx_max) {
		H = (int32_t*)malloc((size_t)tlen_ * 16 * 4);
		for (t = 0; t < tlen_ * 16; ++t) H[t] = NEG_INF;
	}
	if (with_cigar) {
		pm = (uint8_t*)malloc(((size_t)(qlen + tlen - 1) * n_col_ + 1) * 16);
		off = (int*)malloc((size_t)(qlen + tlen - 1) * sizeof(int) * 2);
		off_end = off + qlen + tlen - 1;
	}
	for (t = 0; t < qlen; ++t) qr[t] = query[qlen - 1 - t];
	memcpy(sf, target, tlen);

	for (r = 0, last_st = last_en = -1; r < qlen + tlen - 1; ++r) {
		int st = 0, en = tlen - 1, st0, en0, st_, en_, blk;
		int8_t x1, x21, v1;
		uint8_t *qrr = qr + (qlen - 1 - r);
		if (st < r - qlen + 1) st = r - qlen + 1;
		if (en > r) en = r;
		if (st < (r - w + 1) >> 1) st = (r - w + 1) >> 1;
		if (en > (r + w) >> 1) en = (r + w) >> 1;
		if (st > en) { ez->zdropped = 1; break; }
		st0 = st, en0 = en;
		st = st / 16 * 16, en = (en + 16) / 16 * 16 - 1;
		if (st > 0) { /* ksw2_extd2_sse.c:141-151 */
			if (st - 1 >= last_st && st - 1 <= last_en) x1 = x[st - 1], x21 = x2[st - 1], v1 = v[st - 1];
			else x1 = -q - e, x21 = -q2 - e2, v1 = -q - e;
		} else {
			x1 = -q - e, x21 = -q2 - e2;
			v1 = r == 0 ? -q - e : r < long_thres ? -e : r == long_thres ? long_diff : -e2;
		}
		if (en >= r) {
			y[r] = -q - e, y2[r] = -q2 - e2;
			u[r] = r == 0 ? -q - e : r < long_thres ? -e : r == long_thres ? long_diff : -e2;
		}
		if (!(flag & EZ_GENERIC_SC)) { /* ksw2_extd2_sse.c:157-172: 16-byte chunks from st0, unaligned */
			for (t = st0; t <= en0; t += 16) {
				int l;
				int8_t tmp[16];
				for (l = 0; l < 16; ++l) {
					uint8_t sq = sf[t + l], sq2 = qrr[t + l];
					int8_t z = sq == sq2 ? sc_mch : sc_mis;
					if (sq == (uint8_t)(m - 1) || sq2 == (uint8_t)(m - 1)) z = sc_N;
					tmp[l] = z;
				}
				memcpy(s + t, tmp, 16); /* may spill up to 15 bytes into sf[]: entries below st0, never read again */
			}
		} else {
			for (t = st0; t <= en0; ++t) s[t] = mat[sf[t] * m + qrr[t]];
		}
		st_ = st / 16, en_ = en / 16;
		for (blk = st_; blk <= en_; ++blk) { /* ksw2_extd2_sse.c:184-313 */
			int l, base = blk * 16;
			int8_t nx1 = x[base + 15], nv1 = v[base + 15], nx21 = x2[base + 15];
			int8_t zz[16], aa[16], bb[16], aa2[16], bb2[16], vt1[16], ut[16];
			uint8_t dd[16];
			for (l = 0; l < 16; ++l) { /* __dp_code_block1 */
				int8_t xt1 = l ? x[base + l - 1] : x1;
				int8_t x2t1 = l ? x2[base + l - 1] : x21;
				vt1[l] = l ? v[base + l - 1] : v1;
				ut[l] = u[base + l];
				zz[l] = s[base + l];
				aa[l] = (int8_t)(xt1 + vt1[l]);
				bb[l] = (int8_t)(y[base + l] + ut[l]);
				aa2[l] = (int8_t)(x2t1 + vt1[l]);
				bb2[l] = (int8_t)(y2[base + l] + ut[l]);
			}
			x1 = nx1, v1 = nv1, x21 = nx21;
			for (l = 0; l < 16; ++l) {
				int8_t z = zz[l], a = aa[l], b = bb[l], a2 = aa2[l], b2 = bb2[l], tmp;
				uint8_t d = 0;
				if (!with_cigar) {
					z = z > a ? z : a; z = z > b ? z : b; z = z > a2 ? z : a2; z = z > b2 ? z : b2;
					z = z < sc_mch ? z : sc_mch;
				} else if (!(flag & EZ_RIGHT)) { /* ksw2_extd2_sse.c:227-235 */
					d = a > z ? 1 : 0;   z = z > a ? z : a;
					d = b > z ? 2 : d;   z = z > b ? z : b;
					d = a2 > z ? 3 : d;  z = z > a2 ? z : a2;
					d = b2 > z ? 4 : d;  z = z > b2 ? z : b2;
					z = z < sc_mch ? z : sc_mch;
				} else { /* ksw2_extd2_sse.c:274-282 */
					d = z > a ? 0 : 1;   z = z > a ? z : a;
					d = z > b ? d : 2;   z = z > b ? z : b;
					d = z > a2 ? d : 3;  z = z > a2 ? z : a2;
					d = z > b2 ? d : 4;  z = z > b2 ? z : b2;
					z = z < sc_mch ? z : sc_mch;
				}
				u[base + l] = (int8_t)(z - vt1[l]); /* __dp_code_block2 */
				v[base + l] = (int8_t)(z - ut[l]);
				tmp = (int8_t)(z - q);  a = (int8_t)(a - tmp);  b = (int8_t)(b - tmp);
				tmp = (int8_t)(z - q2); a2 = (int8_t)(a2 - tmp); b2 = (int8_t)(b2 - tmp);
				if (!with_cigar) {
					x[base + l]  = (int8_t)((a  > 0 ? a  : 0) - qe);
					y[base + l]  = (int8_t)((b  > 0 ? b  : 0) - qe);
					x2[base + l] = (int8_t)((a2 > 0 ? a2 : 0) - (q2 + e2));
					y2[base + l] = (int8_t)((b2 > 0 ? b2 : 0) - (q2 + e2));
				} else if (!(flag & EZ_RIGHT)) { /* continuation iff strictly positive */
					x[base + l]  = (int8_t)((a  > 0 ? a  : 0) - qe);        if (a  > 0) d |= 0x08;
					y[base + l]  = (int8_t)((b  > 0 ? b  : 0) - qe);        if (b  > 0) d |= 0x10;
					x2[base + l] = (int8_t)((a2 > 0 ? a2 : 0) - (q2 + e2)); if (a2 > 0) d |= 0x20;
					y2[base + l] = (int8_t)((b2 > 0 ? b2 : 0) - (q2 + e2)); if (b2 > 0) d |= 0x40;
				} else { /* continuation iff non-negative */
					x[base + l]  = (int8_t)((a  < 0 ? 0 : a)  - qe);        if (!(a  < 0)) d |= 0x08;
					y[base + l]  = (int8_t)((b  < 0 ? 0 : b)  - qe);        if (!(b  < 0)) d |= 0x10;
					x2[base + l] = (int8_t)((a2 < 0 ? 0 : a2) - (q2 + e2)); if (!(a2 < 0)) d |= 0x20;
					y2[base + l] = (int8_t)((b2 < 0 ? 0 : b2) - (q2 + e2)); if (!(b2 < 0)) d |= 0x40;
				}
				dd[l] = d;
			}
			if (with_cigar) memcpy(pm + ((size_t)r * n_col_ + (blk - st_)) * 16, dd, 16);
		}
		if (with_cigar) off[r] = st, off_end[r] = en;
		if (!approx_max) { /* ksw2_extd2_sse.c:315-358 */
			int32_t max_H, max_t;
			if (r > 0) {
				int32_t HH[4], tt[4], en1 = st0 + (en0 - st0) / 4 * 4, i;
				max_H = H[en0] = en0 > 0 ? H[en0 - 1] + u[en0] : H[en0] + v[en0];
				max_t = en0;
				for (i = 0; i < 4; ++i) HH[i] = max_H, tt[i] = max_t;
				for (t = st0; t < en1; t += 4)
					for (i = 0; i < 4; ++i) {
						H[t + i] += v[t + i];
						if (H[t + i] > HH[i]) HH[i] = H[t + i], tt[i] = t;
					}
				for (i = 0; i < 4; ++i)
					if (max_H < HH[i]) max_H = HH[i], max_t = tt[i] + i;
				for (; t < en0; ++t) {
					H[t] += (int32_t)v[t];
					if (H[t] > max_H) max_H = H[t], max_t = t;
				}
			} else H[0] = v[0] - qe_pre, max_H = H[0], max_t = 0;
			if (en0 == tlen - 1 && H[en0] > ez->mte) ez->mte = H[en0], ez->mte_q = r - en;
			if (r - st0 == qlen - 1 && H[st0] > ez->mqe) ez->mqe = H[st0], ez->mqe_t = st0;
			if (apply_zdrop(ez, max_H, r, max_t, zdrop, e2)) break;
			if (r == qlen + tlen - 2 && en0 == tlen - 1) ez->score = H[tlen - 1];
		} else { /* ksw2_extd2_sse.c:359-375 */
			if (r > 0) {
				if (last_H0_t >= st0 && last_H0_t <= en0 && last_H0_t + 1 >= st0 && last_H0_t + 1 <= en0) {
					int32_t d0 = v[last_H0_t], d1 = u[last_H0_t + 1];
					if (d0 > d1) H0 += d0;
					else H0 += d1, ++last_H0_t;
				} else if (last_H0_t >= st0 && last_H0_t <= en0) {
					H0 += v[last_H0_t];
				} else {
					++last_H0_t, H0 += u[last_H0_t];
				}
			} else H0 = v[0] - qe_pre, last_H0_t = 0;
			if ((flag & EZ_APPROX_DROP) && apply_zdrop(ez, H0, r, last_H0_t, zdrop, e2)) break;
			if (r == qlen + tlen - 2 && en0 == tlen - 1) ez->score = H0;
		}
		last_st = st, last_en = en;
	}
	free(mem); free(H);
	if (with_cigar) { /* ksw2_extd2_sse.c:381-391 */
		int rev_cigar = !!(flag & EZ_REV_CIGAR);
		if (!ez->zdropped && !(flag & EZ_EXTZ_ONLY)) {
			backtrack(ez, rev_cigar, pm, off, off_end, n_col_ * 16, tlen - 1, qlen - 1);
		} else if (!ez->zdropped && (flag & EZ_EXTZ_ONLY) && ez->mqe + end_bonus > (int)ez->max) {
			ez->reach_end = 1;
			backtrack(ez, rev_cigar, pm, off, off_end, n_col_ * 16, ez->mqe_t, qlen - 1);
		} else if (ez->max_t >= 0 && ez->max_q >= 0) {
			backtrack(ez, rev_cigar, pm, off, off_end, n_col_ * 16, ez->max_t, ez->max_q);
		}
		free(pm); free(off);
	}
}
