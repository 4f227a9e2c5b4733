/* mmsynth -- seeded synthetic reference-pair / read generator for the AirLift re-alignment workloads named in
 * BASELINE.json.  Everything is a pure function of the arguments (xorshift64* streams, one stream per
 * fixed-size block, so the output does not depend on the number of threads).
 *
 *   mmsynth pair <prefix> <genome_bp> <n_contigs> <seed> [edit_seed] [--old]
 *        -> <prefix>.new.fa       the NEW reference (what the reads are re-aligned against)
 *           <prefix>.regions.bed  updated regions in new-reference coordinates (+- 150 bp, merged) -- the intervals AirLift
 *                                 re-aligns reads for (src/2-generate_gaps/2_get_updated_regions_bed.py:94-110)
 *           <prefix>.src.fa       read sources: one window per updated region (+ flanks) and the retired sequences
 *           <prefix>.old.fa       (--old) the OLD reference: the new one with the edits undone
 *   mmsynth srp  <prefix> <out_1.fq> <out_2.fq> <n_pairs> <read_seed>   2x150 FR pairs from <prefix>.src.fa
 *   mmsynth long <ref.fa> <out.fq> <n_reads> <seed> [mean_len [regions.bed]]
 *   mmsynth ref  <out.fa> <genome_bp> <n_contigs> <seed>                the new reference alone
 *   mmsynth sr   <ref.fa> <out_1.fq> <out_2.fq> <n_pairs> <seed> [region_frac [read_seed]]   (small test inputs)
 *
 * new reference: uniform random ACGT in n_contigs contigs "chr1..", plus repeat families -- one 2 kb family x50 copies per
 *     2 Mbp, a 300 bp family x G/31000 copies (10^5 at human size), a 6 kb family x G/310000 copies (10^4 at human size),
 *     every copy mutated by its own 1-10 % -- and ~0.1 % of the bases inside short N runs.
 * edits (new -> old, edit_seed): SNPs 1e-4 (no region: chain blocks run across mismatches), small indels 1e-5 (1-20 bp),
 *     and clamp(G/310000, 100, 10000) segments of 100 bp..100 kb (log-uniform) that are new-only (absent from the old
 *     reference), retired (old-only sequence at a breakpoint) or moved (same sequence elsewhere in the old reference).
 * reads: insert ~N(400,50) clipped to [250,800], 50/50 strand, 1 % substitutions, 0.15 % insertions, 0.05 % deletions;
 *     97 % of the fragments overlap an updated region, 3 % come from retired sequence with its flanks.
 * long : ONT-shaped reads, log-normal length (mean mean_len, clipped [1k,100k]), 3 % sub + 3 % ins + 3 % del; with a BED,
 *     every read overlaps an updated region.
 */
#define _GNU_SOURCE
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
#include <math.h>
#include <pthread.h>
#include <unistd.h>

typedef struct { uint64_t s; } rng_t;
static inline uint64_t rng_next(rng_t *r)
{
	uint64_t x = r->s;
	x ^= x >> 12; x ^= x << 25; x ^= x >> 27;
	r->s = x;
	return x * 0x2545F4914F6CDD1DULL;
}
static inline void rng_seed(rng_t *r, uint64_t seed)
{
	r->s = seed * 0x9E3779B97F4A7C15ULL + 0xD1B54A32D192ED03ULL;
	if (r->s == 0) r->s = 1;
	for (int i = 0; i < 8; ++i) rng_next(r);
}
static inline double rng_unif(rng_t *r) { return (rng_next(r) >> 11) * (1.0 / 9007199254740992.0); }
static inline uint64_t rng_below(rng_t *r, uint64_t n) { return (uint64_t)(rng_unif(r) * (double)n); }
static inline double rng_norm(rng_t *r)
{
	double u1 = rng_unif(r), u2 = rng_unif(r);
	if (u1 < 1e-300) u1 = 1e-300;
	return sqrt(-2.0 * log(u1)) * cos(6.283185307179586 * u2);
}
static inline int64_t rng_loguni(rng_t *r, double lo, double hi) { return (int64_t)exp(log(lo) + rng_unif(r) * (log(hi) - log(lo))); }

static const char BASES[5] = "ACGT";
static inline char comp(char c)
{
	switch (c) { case 'A': return 'T'; case 'C': return 'G'; case 'G': return 'C'; case 'T': return 'A'; default: return 'N'; }
}

static int n_workers(void)
{
	long n = sysconf(_SC_NPROCESSORS_ONLN);
	const char *e = getenv("MMSYNTH_THREADS");
	if (e && atoi(e) > 0) n = atoi(e);
	return n < 1 ? 1 : n > 16 ? 16 : (int)n;
}

/* ---- block-parallel helper: fn(block index) over [0, n_blocks), blocks handed out by an atomic counter */
typedef struct { void (*fn)(void*, int64_t); void *arg; int64_t n, next; pthread_mutex_t mu; } par_t;
static void *par_main(void *p_)
{
	par_t *p = (par_t*)p_;
	for (;;) {
		pthread_mutex_lock(&p->mu);
		const int64_t i = p->next++;
		pthread_mutex_unlock(&p->mu);
		if (i >= p->n) return 0;
		p->fn(p->arg, i);
	}
}
static void par_for(void (*fn)(void*, int64_t), void *arg, int64_t n)
{
	par_t p = {fn, arg, n, 0, PTHREAD_MUTEX_INITIALIZER};
	pthread_t th[16];
	int nt = n_workers(), i;
	if (nt > n) nt = (int)n;
	for (i = 1; i < nt; ++i) pthread_create(&th[i], 0, par_main, &p);
	par_main(&p);
	for (i = 1; i < nt; ++i) pthread_join(th[i], 0);
}

/* ------------------------------------------------------------------ new reference */
#define GEN_BLOCK (1 << 22)
typedef struct { char *g; int64_t G; uint64_t seed; } fill_t;
static void fill_block(void *a_, int64_t b)
{
	fill_t *a = (fill_t*)a_;
	rng_t r; rng_seed(&r, a->seed ^ (0x51ED270B1ULL * (uint64_t)(b + 1)));
	int64_t i = b * (int64_t)GEN_BLOCK, e = i + GEN_BLOCK < a->G ? i + GEN_BLOCK : a->G;
	while (i < e) {
		uint64_t x = rng_next(&r);
		for (int j = 0; j < 32 && i < e; ++j, x >>= 2) a->g[i++] = BASES[x & 3];
	}
}

static void plant_family(rng_t *r, char *g, int64_t G, int blk_len, int64_t copies)
{
	if (G <= blk_len + 1) return;
	int64_t src = rng_below(r, G - blk_len);
	char *blk = (char*)malloc(blk_len);
	memcpy(blk, g + src, blk_len);
	for (int64_t c = 0; c < copies; ++c) {
		int64_t dst = rng_below(r, G - blk_len);
		double div = 0.01 + 0.09 * rng_unif(r);
		/* geometric skipping: mutate every ~1/div-th base */
		memcpy(g + dst, blk, blk_len);
		for (double at = log(1.0 - rng_unif(r)) / log(1.0 - div); at < blk_len; at += 1.0 + log(1.0 - rng_unif(r)) / log(1.0 - div))
			g[dst + (int)at] = BASES[rng_next(r) >> 62];
	}
	free(blk);
}

static char *make_genome(int64_t G, uint64_t seed)
{
	char *g = (char*)malloc(G + 1);
	if (!g) return 0;
	fill_t f = {g, G, seed};
	par_for(fill_block, &f, (G + GEN_BLOCK - 1) / GEN_BLOCK);
	g[G] = 0;
	rng_t r; rng_seed(&r, seed);
	int64_t n_fam = G / 2000000; if (n_fam < 1) n_fam = 1;
	for (int64_t i = 0; i < n_fam; ++i) plant_family(&r, g, G, 2000, 50);
	plant_family(&r, g, G, 300, G / 31000);     /* short high-copy family: beyond max_occ, shows up as rep_len */
	plant_family(&r, g, G, 6000, G / 310000);   /* long family: between mid_occ and max_occ at human size -> re-chain path */
	for (int64_t covered = 0; covered < G / 1000; ) { /* N runs: ~0.1 % of bases, runs of 10-200 */
		int64_t len = 10 + rng_below(&r, 191), at = rng_below(&r, G - len);
		memset(g + at, 'N', len);
		covered += len;
	}
	return g;
}

typedef struct { int n; int64_t *off, *len; } ctg_t;
static ctg_t split_contigs(int64_t G, int n_ctg)
{
	ctg_t c; c.n = n_ctg; c.off = (int64_t*)malloc((n_ctg + 1) * 8); c.len = (int64_t*)malloc(n_ctg * 8);
	int64_t per = G / n_ctg, off = 0;
	for (int i = 0; i < n_ctg; ++i) { c.off[i] = off; c.len[i] = i == n_ctg - 1 ? G - off : per; off += c.len[i]; }
	c.off[n_ctg] = G;
	return c;
}

static void write_seq(FILE *fp, const char *s, int64_t len)
{ /* 60 bases per line, formatted in 1 MB chunks */
	enum { LINES = 16384 };
	char *buf = (char*)malloc((size_t)LINES * 61);
	for (int64_t i = 0; i < len; ) {
		size_t o = 0;
		for (int l = 0; l < LINES && i < len; ++l) {
			int64_t n = len - i < 60 ? len - i : 60;
			memcpy(buf + o, s + i, n); o += n; buf[o++] = '\n'; i += n;
		}
		fwrite(buf, 1, o, fp);
	}
	free(buf);
}

static int write_fasta(const char *fn, const char *g, const ctg_t *c)
{
	FILE *fp = fopen(fn, "w");
	if (!fp) return 1;
	setvbuf(fp, 0, _IOFBF, 8 << 20);
	for (int i = 0; i < c->n; ++i) { fprintf(fp, ">chr%d\n", i + 1); write_seq(fp, g + c->off[i], c->len[i]); }
	return fclose(fp) != 0;
}

/* ------------------------------------------------------------------ edits: the reference pair */
enum { E_SNP, E_INS_NEW, E_DEL_NEW, E_SEG_NEW, E_SEG_RETIRED, E_SEG_MOVED };
typedef struct { int64_t pos, len; int kind; int64_t aux; } edit_t; /* pos in concatenated new coordinates; aux: old-side anchor of a moved segment */
static int edit_cmp(const void *a, const void *b) { int64_t x = ((const edit_t*)a)->pos, y = ((const edit_t*)b)->pos; return x < y ? -1 : x > y; }

static int ctg_of(const ctg_t *c, int64_t p) { int lo = 0, hi = c->n - 1; while (lo < hi) { int m = (lo + hi + 1) >> 1; if (c->off[m] <= p) lo = m; else hi = m - 1; } return lo; }

static void retired_seq(uint64_t edit_seed, int64_t idx, int64_t len, char *out)
{
	rng_t r; rng_seed(&r, edit_seed ^ (0xA24BAED4963EE407ULL * (uint64_t)(idx + 1)));
	for (int64_t i = 0; i < len; ) { uint64_t x = rng_next(&r); for (int j = 0; j < 32 && i < len; ++j, x >>= 2) out[i++] = BASES[x & 3]; }
}

typedef struct { int ctg; int64_t st, en; } region_t;
static int region_cmp(const void *a, const void *b)
{
	const region_t *x = (const region_t*)a, *y = (const region_t*)b;
	if (x->ctg != y->ctg) return x->ctg < y->ctg ? -1 : 1;
	return x->st < y->st ? -1 : x->st > y->st;
}

#define MARGIN 150      /* read length: 2_get_updated_regions_bed.py extends each gap by read_size (+ max_errors) */
#define FLANK  800      /* so that a whole fragment overlapping a region lies inside its window */

static int gen_pair(const char *prefix, int64_t G, int n_ctg, uint64_t seed, uint64_t edit_seed, int want_old)
{
	char fn[4096];
	char *g = make_genome(G, seed);
	if (!g) { fprintf(stderr, "mmsynth: out of memory\n"); return 1; }
	ctg_t c = split_contigs(G, n_ctg);
	snprintf(fn, sizeof fn, "%s.new.fa", prefix);
	if (write_fasta(fn, g, &c)) return 1;

	rng_t r; rng_seed(&r, edit_seed);
	int64_t n_seg = G / 310000; if (n_seg < 100) n_seg = 100; if (n_seg > 10000) n_seg = 10000;
	int64_t n_indel = G / 100000, n_snp = want_old ? G / 10000 : 0;
	int64_t n_e = 0, m_e = n_seg + n_indel + n_snp + 16;
	edit_t *e = (edit_t*)malloc(m_e * sizeof(edit_t));
	for (int64_t i = 0; i < n_seg; ++i) {
		edit_t x; double u = rng_unif(&r);
		x.kind = u < 0.4 ? E_SEG_NEW : u < 0.7 ? E_SEG_RETIRED : E_SEG_MOVED;
		x.len = rng_loguni(&r, 100.0, 100000.0);
		x.pos = rng_below(&r, G); x.aux = rng_below(&r, G);
		e[n_e++] = x;
	}
	for (int64_t i = 0; i < n_indel; ++i) {
		edit_t x; x.kind = (rng_next(&r) >> 63) ? E_INS_NEW : E_DEL_NEW; x.len = 1 + rng_below(&r, 20); x.pos = rng_below(&r, G); x.aux = 0;
		e[n_e++] = x;
	}
	for (int64_t i = 0; i < n_snp; ++i) { edit_t x; x.kind = E_SNP; x.len = 1; x.pos = rng_below(&r, G); x.aux = 0; e[n_e++] = x; }
	qsort(e, n_e, sizeof(edit_t), edit_cmp);
	/* keep edits that lie inside one contig, away from its ends, and do not touch the previous edit */
	int64_t k = 0, last_end = -1;
	for (int64_t i = 0; i < n_e; ++i) {
		edit_t x = e[i];
		const int ci = ctg_of(&c, x.pos);
		const int64_t span = (x.kind == E_SEG_RETIRED || x.kind == E_DEL_NEW) ? 0 : x.len;
		if (x.pos - c.off[ci] < 2000 || x.pos + span > c.off[ci + 1] - 2000) continue;
		if (x.pos <= last_end + 1) continue;
		e[k++] = x; last_end = x.pos + span;
	}
	n_e = k;

	/* regions (new coordinates) */
	region_t *reg = (region_t*)malloc((n_e + 1) * sizeof(region_t)); int64_t n_reg = 0;
	for (int64_t i = 0; i < n_e; ++i) {
		if (e[i].kind == E_SNP) continue;
		const int ci = ctg_of(&c, e[i].pos);
		const int64_t span = (e[i].kind == E_SEG_RETIRED || e[i].kind == E_DEL_NEW) ? 0 : e[i].len;
		int64_t st = e[i].pos - c.off[ci] - MARGIN, en = e[i].pos - c.off[ci] + span + MARGIN;
		if (st < 0) st = 0; if (en > c.len[ci]) en = c.len[ci];
		reg[n_reg].ctg = ci, reg[n_reg].st = st, reg[n_reg].en = en; ++n_reg;
	}
	qsort(reg, n_reg, sizeof(region_t), region_cmp);
	k = 0;
	for (int64_t i = 0; i < n_reg; ++i) {
		if (k && reg[k-1].ctg == reg[i].ctg && reg[i].st <= reg[k-1].en) { if (reg[i].en > reg[k-1].en) reg[k-1].en = reg[i].en; }
		else reg[k++] = reg[i];
	}
	n_reg = k;
	snprintf(fn, sizeof fn, "%s.regions.bed", prefix);
	FILE *fb = fopen(fn, "w");
	snprintf(fn, sizeof fn, "%s.src.fa", prefix);
	FILE *fs = fopen(fn, "w");
	if (!fb || !fs) return 1;
	setvbuf(fs, 0, _IOFBF, 8 << 20);
	int64_t cov = 0;
	for (int64_t i = 0; i < n_reg; ++i) {
		const int ci = reg[i].ctg;
		int64_t ws = reg[i].st - FLANK, we = reg[i].en + FLANK;
		if (ws < 0) ws = 0; if (we > c.len[ci]) we = c.len[ci];
		fprintf(fb, "chr%d\t%ld\t%ld\n", ci + 1, (long)reg[i].st, (long)reg[i].en);
		/* header: window start, then the region inside the window */
		fprintf(fs, ">chr%d:%ld:%ld:%ld\n", ci + 1, (long)ws, (long)(reg[i].st - ws), (long)(reg[i].en - ws));
		write_seq(fs, g + c.off[ci] + ws, we - ws);
		cov += reg[i].en - reg[i].st;
	}
	fclose(fb);
	/* retired sequences with their new-reference flanks (reads that straddle the breakpoint are half-mappable) */
	int64_t n_ret = 0;
	for (int64_t i = 0; i < n_e; ++i) {
		if (e[i].kind != E_SEG_RETIRED) continue;
		const int ci = ctg_of(&c, e[i].pos);
		int64_t ls = e[i].pos - FLANK < c.off[ci] ? c.off[ci] : e[i].pos - FLANK, re = e[i].pos + FLANK > c.off[ci + 1] ? c.off[ci + 1] : e[i].pos + FLANK;
		char *buf = (char*)malloc((e[i].pos - ls) + e[i].len + (re - e[i].pos) + 1);
		memcpy(buf, g + ls, e[i].pos - ls);
		retired_seq(edit_seed, i, e[i].len, buf + (e[i].pos - ls));
		memcpy(buf + (e[i].pos - ls) + e[i].len, g + e[i].pos, re - e[i].pos);
		fprintf(fs, ">retired%ld:%ld:%ld:%ld\n", (long)n_ret, 0L, (long)(e[i].pos - ls), (long)(e[i].pos - ls + e[i].len));
		write_seq(fs, buf, (e[i].pos - ls) + e[i].len + (re - e[i].pos));
		free(buf); ++n_ret;
	}
	fclose(fs);
	fprintf(stderr, "[mmsynth] %ld bp in %d contigs; %ld edits -> %ld updated regions covering %ld bp (%.2f %%), %ld retired segments\n",
			(long)G, n_ctg, (long)n_e, (long)n_reg, (long)cov, 100.0 * cov / G, (long)n_ret);

	if (want_old) { /* undo the edits, left to right */
		snprintf(fn, sizeof fn, "%s.old.fa", prefix);
		FILE *fo = fopen(fn, "w");
		if (!fo) return 1;
		setvbuf(fo, 0, _IOFBF, 8 << 20);
		int64_t ei = 0;
		/* moved segments re-appear at their old anchor: collect (anchor, source) pairs sorted by anchor */
		int64_t n_mv = 0; edit_t *mv = (edit_t*)malloc((n_e + 1) * sizeof(edit_t));
		for (int64_t i = 0; i < n_e; ++i) if (e[i].kind == E_SEG_MOVED) { mv[n_mv].pos = e[i].aux, mv[n_mv].len = e[i].len, mv[n_mv].aux = e[i].pos, mv[n_mv].kind = 0; ++n_mv; }
		qsort(mv, n_mv, sizeof(edit_t), edit_cmp);
		int64_t mi = 0;
		for (int ci = 0; ci < c.n; ++ci) {
			size_t cap = (size_t)c.len[ci] + (1 << 20), o = 0;
			char *buf = (char*)malloc(cap);
#define PUT(src, n) do { if (o + (size_t)(n) + 1 > cap) { cap = (o + (size_t)(n)) * 3 / 2 + 1024; buf = (char*)realloc(buf, cap); } memcpy(buf + o, (src), (n)); o += (n); } while (0)
			int64_t p = c.off[ci];
			const int64_t end = c.off[ci + 1];
			while (p < end) {
				int64_t next_e = ei < n_e && e[ei].pos < end ? e[ei].pos : end;
				int64_t next_m = mi < n_mv && mv[mi].pos < end ? mv[mi].pos : end;
				if (next_m < p) next_m = p;
				if (next_m < next_e) { /* a moved segment's old home lies here (never inside another edit: it lands between edits) */
					PUT(g + p, next_m - p); p = next_m;
					PUT(g + mv[mi].aux, mv[mi].len); ++mi;
					continue;
				}
				PUT(g + p, next_e - p); p = next_e;
				if (p >= end) break;
				const edit_t x = e[ei++];
				switch (x.kind) {
				case E_SNP: { char b = g[p]; if (b != 'N') { char n; rng_t q; rng_seed(&q, edit_seed ^ (uint64_t)p); do n = BASES[rng_next(&q) >> 62]; while (n == b); b = n; } PUT(&b, 1); p += 1; break; }
				case E_INS_NEW: case E_SEG_NEW: case E_SEG_MOVED: p += x.len; break; /* absent from the old reference at this place */
				case E_DEL_NEW: case E_SEG_RETIRED: { char *t = (char*)malloc(x.len); retired_seq(edit_seed, ei - 1, x.len, t); PUT(t, x.len); free(t); break; }
				}
			}
			fprintf(fo, ">chr%d\n", ci + 1);
			write_seq(fo, buf, (int64_t)o);
			free(buf);
#undef PUT
		}
		free(mv);
		fclose(fo);
	}
	free(e); free(reg); free(g);
	return 0;
}

static int gen_ref(const char *fn, int64_t G, int n_ctg, uint64_t seed)
{
	char *g = make_genome(G, seed);
	if (!g) return 1;
	ctg_t c = split_contigs(G, n_ctg);
	int rc = write_fasta(fn, g, &c);
	free(g);
	return rc;
}

/* ------------------------------------------------------------------ FASTA loading */
typedef struct { int n; char **name; char **seq; int64_t *len; int64_t total; } ref_t;

static ref_t *load_ref(const char *fn)
{
	FILE *fp = fopen(fn, "r");
	if (!fp) return 0;
	setvbuf(fp, 0, _IOFBF, 8 << 20);
	ref_t *R = (ref_t*)calloc(1, sizeof(ref_t));
	size_t cap = 0; char *line = 0; ssize_t l;
	int m = 0; int64_t cur_m = 0;
	while ((l = getline(&line, &cap, fp)) > 0) {
		while (l > 0 && (line[l-1] == '\n' || line[l-1] == '\r')) line[--l] = 0;
		if (line[0] == '>') {
			if (R->n == m) {
				m = m ? m * 2 : 16;
				R->name = (char**)realloc(R->name, m * sizeof(char*));
				R->seq = (char**)realloc(R->seq, m * sizeof(char*));
				R->len = (int64_t*)realloc(R->len, m * sizeof(int64_t));
			}
			char *p = line + 1; while (*p && *p != ' ' && *p != '\t') ++p; *p = 0;
			R->name[R->n] = strdup(line + 1); R->seq[R->n] = 0; R->len[R->n] = 0; cur_m = 0; ++R->n;
		} else if (R->n > 0) {
			int i = R->n - 1;
			if (R->len[i] + l + 1 > cur_m) { cur_m = (R->len[i] + l + 1) * 2; R->seq[i] = (char*)realloc(R->seq[i], cur_m); }
			memcpy(R->seq[i] + R->len[i], line, l); R->len[i] += l; R->seq[i][R->len[i]] = 0;
		}
	}
	free(line); fclose(fp);
	for (int i = 0; i < R->n; ++i) R->total += R->len[i];
	return R;
}

/* mutate src[0..len) into dst (capacity cap) with the given error rates; returns new length */
static int mutate(rng_t *r, const char *src, int len, char *dst, int cap, double sub, double ins, double del, int want)
{
	int o = 0;
	for (int i = 0; i < len && o < cap && (want <= 0 || o < want); ++i) {
		double u = rng_unif(r);
		if (u < del) continue;
		if (u < del + ins) { dst[o++] = BASES[rng_next(r) >> 62]; if (o >= cap) break; }
		if (u < del + ins + sub && src[i] != 'N') {
			char c; do c = BASES[rng_next(r) >> 62]; while (c == src[i]);
			dst[o++] = c;
		} else dst[o++] = src[i];
	}
	return o;
}

/* the same error model, jumping from error to error (errors are ~1 % of the bases) */
static int mutate_fast(rng_t *r, const char *src, int len, char *dst, int want, double sub, double ins, double del)
{
	const double p = sub + ins + del, lq = log(1.0 - p);
	int i = 0, o = 0;
	while (i < len && o < want) {
		int run = (int)(log(1.0 - rng_unif(r)) / lq); /* error-free bases before the next error */
		if (run > len - i) run = len - i;
		if (run > want - o) run = want - o;
		memcpy(dst + o, src + i, run); o += run; i += run;
		if (i >= len || o >= want) break;
		const double u = rng_unif(r) * p;
		if (u < del) { ++i; continue; }
		if (u < del + ins) { dst[o++] = BASES[rng_next(r) >> 62]; continue; }
		if (src[i] != 'N') { char c; do c = BASES[rng_next(r) >> 62]; while (c == src[i]); dst[o++] = c; } else dst[o++] = 'N';
		++i;
	}
	return o;
}

static void revcomp(char *s, int l)
{
	for (int i = 0; i < l >> 1; ++i) { char t = comp(s[i]); s[i] = comp(s[l-1-i]); s[l-1-i] = t; }
	if (l & 1) s[l>>1] = comp(s[l>>1]);
}

/* ------------------------------------------------------------------ short reads from a pair's source windows */
#define SR_BLOCK 16384
typedef struct {
	ref_t *R; int n_upd, n_ret; double *cum; double tot; int64_t *rst, *ren; /* region inside each window */
	uint64_t read_seed; int64_t n_pairs, b0; /* b0: first block of the group being generated */
	char **out1, **out2; size_t *len1, *len2;
} srp_t;

static void srp_block(void *a_, int64_t b)
{
	srp_t *a = (srp_t*)a_;
	b += a->b0;
	rng_t r; rng_seed(&r, a->read_seed ^ (0x9FB21C651E98DF25ULL * (uint64_t)(b + 1)));
	const int64_t p0 = b * SR_BLOCK, p1 = p0 + SR_BLOCK < a->n_pairs ? p0 + SR_BLOCK : a->n_pairs;
	char *o1 = (char*)malloc((size_t)(p1 - p0) * 360), *o2 = (char*)malloc((size_t)(p1 - p0) * 360);
	size_t l1 = 0, l2 = 0;
	char q[160]; memset(q, 'I', 150);
	char frag[1024], tmp[1024], m1[160], m2[160];
	for (int64_t p = p0; p < p1; ++p) {
		int ins = (int)(400.0 + 50.0 * rng_norm(&r) + 0.5);
		if (ins < 250) ins = 250; if (ins > 800) ins = 800;
		int w; int64_t st;
		if (a->n_ret > 0 && rng_unif(&r) < 0.03) w = a->n_upd + (int)rng_below(&r, a->n_ret);
		else {
			double u = rng_unif(&r) * a->tot; int lo = 0, hi = a->n_upd - 1;
			while (lo < hi) { int mid = (lo + hi) >> 1; if (a->cum[mid] < u) lo = mid + 1; else hi = mid; }
			w = lo;
		}
		/* fragment start such that the fragment overlaps [rst, ren) of the window */
		{
			int64_t lo = a->rst[w] - ins + 1, hi = a->ren[w] - 1, wl = a->R->len[w];
			if (lo < 0) lo = 0;
			if (hi > wl - ins) hi = wl - ins;
			if (hi < lo) { lo = 0; hi = wl - ins; if (hi < 0) { hi = 0; ins = (int)wl; } }
			st = lo + (int64_t)rng_below(&r, hi - lo + 1);
		}
		memcpy(frag, a->R->seq[w] + st, ins);
		if (rng_next(&r) >> 63) revcomp(frag, ins);
		int n1 = mutate_fast(&r, frag, ins < 170 ? ins : 170, m1, 150, 0.01, 0.0015, 0.0005);
		memcpy(tmp, frag, ins); revcomp(tmp, ins);
		int n2 = mutate_fast(&r, tmp, ins < 170 ? ins : 170, m2, 150, 0.01, 0.0015, 0.0005);
		l1 += sprintf(o1 + l1, "@read%ld/1\n", (long)p); memcpy(o1 + l1, m1, n1); l1 += n1; memcpy(o1 + l1, "\n+\n", 3); l1 += 3; memcpy(o1 + l1, q, n1); l1 += n1; o1[l1++] = '\n';
		l2 += sprintf(o2 + l2, "@read%ld/2\n", (long)p); memcpy(o2 + l2, m2, n2); l2 += n2; memcpy(o2 + l2, "\n+\n", 3); l2 += 3; memcpy(o2 + l2, q, n2); l2 += n2; o2[l2++] = '\n';
	}
	a->out1[b] = o1, a->out2[b] = o2, a->len1[b] = l1, a->len2[b] = l2;
}

static int gen_srp(const char *prefix, const char *fn1, const char *fn2, int64_t n_pairs, uint64_t read_seed)
{
	char fn[4096];
	snprintf(fn, sizeof fn, "%s.src.fa", prefix);
	ref_t *R = load_ref(fn);
	if (!R || R->n == 0) { fprintf(stderr, "mmsynth: cannot read %s\n", fn); return 1; }
	srp_t a; memset(&a, 0, sizeof a);
	a.R = R; a.read_seed = read_seed; a.n_pairs = n_pairs;
	a.rst = (int64_t*)malloc(R->n * 8); a.ren = (int64_t*)malloc(R->n * 8); a.cum = (double*)malloc(R->n * sizeof(double));
	for (int i = 0; i < R->n; ++i) {
		long ws, rs, re; char *p = strchr(R->name[i], ':');
		if (!p || sscanf(p + 1, "%ld:%ld:%ld", &ws, &rs, &re) != 3) { fprintf(stderr, "mmsynth: bad window header '%s'\n", R->name[i]); return 1; }
		a.rst[i] = rs, a.ren[i] = re;
		if (strncmp(R->name[i], "retired", 7) == 0) ++a.n_ret;
		else { if (a.n_ret) { fprintf(stderr, "mmsynth: windows out of order\n"); return 1; } ++a.n_upd; a.tot += (double)(re - rs + 400); a.cum[i] = a.tot; }
	}
	if (a.n_upd == 0) { fprintf(stderr, "mmsynth: no updated region in %s\n", fn); return 1; }
	const int64_t n_blk = (n_pairs + SR_BLOCK - 1) / SR_BLOCK;
	FILE *f1 = fopen(fn1, "w"), *f2 = fopen(fn2, "w");
	if (!f1 || !f2) return 1;
	/* groups of blocks generated in parallel, written in order */
	const int64_t grp = 64;
	a.out1 = (char**)calloc(n_blk, sizeof(char*)); a.out2 = (char**)calloc(n_blk, sizeof(char*));
	a.len1 = (size_t*)calloc(n_blk, sizeof(size_t)); a.len2 = (size_t*)calloc(n_blk, sizeof(size_t));
	for (int64_t b0 = 0; b0 < n_blk; b0 += grp) {
		const int64_t nb = n_blk - b0 < grp ? n_blk - b0 : grp;
		a.b0 = b0;
		par_for(srp_block, &a, nb);
		for (int64_t b = b0; b < b0 + nb; ++b) {
			fwrite(a.out1[b], 1, a.len1[b], f1); fwrite(a.out2[b], 1, a.len2[b], f2);
			free(a.out1[b]); free(a.out2[b]);
		}
	}
	return (fclose(f1) != 0) | (fclose(f2) != 0);
}

/* ------------------------------------------------------------------ small-input generators (tests, golden fixtures) */
static int gen_sr(const char *ref_fn, const char *fn1, const char *fn2, int64_t n_pairs, uint64_t seed, double region_frac, uint64_t read_seed)
{
	ref_t *R = load_ref(ref_fn);
	if (!R) return 1;
	rng_t r; rng_seed(&r, seed);
	/* updated regions: intervals of 100 bp..100 kb (log-uniform) until region_frac of genome covered */
	int n_reg = 0, m_reg = 0; region_t *reg = 0; int64_t cov = 0;
	while (cov < (int64_t)(R->total * region_frac) || n_reg == 0) {
		int c; do c = (int)rng_below(&r, R->n); while (R->len[c] < 2000);
		int64_t len = (int64_t)exp(log(100.0) + rng_unif(&r) * (log(100000.0) - log(100.0)));
		if (len > R->len[c] - 1700) len = R->len[c] - 1700;
		int64_t st = 800 + rng_below(&r, R->len[c] - 1600 - len);
		if (n_reg == m_reg) { m_reg = m_reg ? m_reg * 2 : 256; reg = (region_t*)realloc(reg, m_reg * sizeof(region_t)); }
		reg[n_reg].ctg = c; reg[n_reg].st = st; reg[n_reg].en = st + len; ++n_reg; cov += len;
	}
	/* the updated regions belong to the reference pair (seed); shards of one job draw different reads (read_seed) from them */
	if (read_seed && read_seed != seed) rng_seed(&r, read_seed);
	/* cumulative lengths for sampling a region proportional to (len + 2*150) */
	double *cum = (double*)malloc(n_reg * sizeof(double)); double tot = 0;
	for (int i = 0; i < n_reg; ++i) { tot += (double)(reg[i].en - reg[i].st + 300); cum[i] = tot; }
	FILE *f1 = fopen(fn1, "w"), *f2 = fopen(fn2, "w");
	if (!f1 || !f2) return 1;
	char q[160]; memset(q, 'I', 150); q[150] = 0;
	char frag[1024], m1[256], m2[256], tmp[1024];
	for (int64_t p = 0; p < n_pairs; ++p) {
		int ins = (int)(400.0 + 50.0 * rng_norm(&r) + 0.5);
		if (ins < 250) ins = 250; if (ins > 800) ins = 800;
		if (rng_unif(&r) < 0.03) { /* retired sequence: absent from the new reference */
			for (int i = 0; i < ins; ++i) frag[i] = BASES[rng_next(&r) >> 62];
		} else {
			double u = rng_unif(&r) * tot; int lo = 0, hi = n_reg - 1;
			while (lo < hi) { int mid = (lo + hi) >> 1; if (cum[mid] < u) lo = mid + 1; else hi = mid; }
			region_t *g = &reg[lo];
			int64_t span = g->en - g->st + 300 - 150; if (span < 1) span = 1;
			int64_t st = g->st - 150 + (int64_t)rng_below(&r, span) - (ins - 150) / 2;
			if (st < 0) st = 0;
			if (st + ins > R->len[g->ctg]) st = R->len[g->ctg] - ins;
			memcpy(frag, R->seq[g->ctg] + st, ins);
		}
		if (rng_next(&r) >> 63) revcomp(frag, ins);
		/* mate 1 = first 150 of fragment, mate 2 = revcomp of last 150 (FR) */
		int l1 = mutate(&r, frag, ins < 170 ? ins : 170, m1, 150, 0.01, 0.0015, 0.0005, 150);
		memcpy(tmp, frag, ins); revcomp(tmp, ins);
		int l2 = mutate(&r, tmp, ins < 170 ? ins : 170, m2, 150, 0.01, 0.0015, 0.0005, 150);
		m1[l1] = 0; m2[l2] = 0;
		fprintf(f1, "@read%ld/1\n%s\n+\n%.*s\n", (long)p, m1, l1, q);
		fprintf(f2, "@read%ld/2\n%s\n+\n%.*s\n", (long)p, m2, l2, q);
	}
	fclose(f1); fclose(f2);
	return 0;
}

static int gen_long(const char *ref_fn, const char *fn, int64_t n, uint64_t seed, double mean_len, const char *bed_fn)
{
	ref_t *R = load_ref(ref_fn);
	if (!R) return 1;
	rng_t r; rng_seed(&r, seed);
	FILE *fp = fopen(fn, "w");
	if (!fp) return 1;
	setvbuf(fp, 0, _IOFBF, 8 << 20);
	/* optional: every read overlaps an updated region of the BED */
	region_t *reg = 0; int64_t n_reg = 0, m_reg = 0;
	if (bed_fn) {
		FILE *fb = fopen(bed_fn, "r"); char nm[256]; long st, en;
		if (!fb) return 1;
		while (fscanf(fb, "%255s %ld %ld", nm, &st, &en) == 3) {
			int c; for (c = 0; c < R->n; ++c) if (strcmp(R->name[c], nm) == 0) break;
			if (c == R->n) continue;
			if (n_reg == m_reg) { m_reg = m_reg ? m_reg * 2 : 256; reg = (region_t*)realloc(reg, m_reg * sizeof(region_t)); }
			reg[n_reg].ctg = c, reg[n_reg].st = st, reg[n_reg].en = en; ++n_reg;
		}
		fclose(fb);
		if (n_reg == 0) { fprintf(stderr, "mmsynth: no usable region in %s\n", bed_fn); return 1; }
	}
	double sigma = 0.6, mu = log(mean_len) - 0.5 * sigma * sigma;
	int cap = 140000; char *out = (char*)malloc(cap + 1), *src = (char*)malloc(110000), *qual = (char*)malloc(cap + 1);
	memset(qual, '5', cap);
	for (int64_t i = 0; i < n; ++i) {
		int len = (int)exp(mu + sigma * rng_norm(&r));
		if (len < 1000) len = 1000; if (len > 100000) len = 100000;
		int c; int64_t st;
		if (n_reg) {
			const region_t *g = &reg[rng_below(&r, n_reg)];
			c = g->ctg;
			if (len > R->len[c]) len = (int)R->len[c];
			int64_t lo = g->st - len + 1, hi = g->en - 1;
			if (lo < 0) lo = 0; if (hi > R->len[c] - len) hi = R->len[c] - len; if (hi < lo) hi = lo;
			st = lo + (int64_t)rng_below(&r, hi - lo + 1);
		} else {
			do c = (int)rng_below(&r, R->n); while (R->len[c] < 2000);
			if (len > R->len[c]) len = (int)R->len[c];
			st = rng_below(&r, R->len[c] - len + 1);
		}
		memcpy(src, R->seq[c] + st, len);
		if (rng_next(&r) >> 63) revcomp(src, len);
		int l = mutate(&r, src, len, out, cap, 0.03, 0.03, 0.03, 0);
		out[l] = 0;
		fprintf(fp, "@long%ld\n%s\n+\n%.*s\n", (long)i, out, l, qual);
	}
	fclose(fp);
	return 0;
}

int main(int argc, char **argv)
{
	if (argc >= 6 && strcmp(argv[1], "pair") == 0) {
		int want_old = 0; uint64_t edit_seed = 43;
		for (int i = 6; i < argc; ++i) { if (strcmp(argv[i], "--old") == 0) want_old = 1; else edit_seed = strtoull(argv[i], 0, 10); }
		return gen_pair(argv[2], atoll(argv[3]), atoi(argv[4]), strtoull(argv[5], 0, 10), edit_seed, want_old);
	}
	if (argc >= 7 && strcmp(argv[1], "srp") == 0)
		return gen_srp(argv[2], argv[3], argv[4], atoll(argv[5]), strtoull(argv[6], 0, 10));
	if (argc >= 6 && strcmp(argv[1], "ref") == 0)
		return gen_ref(argv[2], atoll(argv[3]), atoi(argv[4]), strtoull(argv[5], 0, 10));
	if (argc >= 7 && strcmp(argv[1], "sr") == 0)
		return gen_sr(argv[2], argv[3], argv[4], atoll(argv[5]), strtoull(argv[6], 0, 10), argc > 7 ? atof(argv[7]) : 0.02, argc > 8 ? strtoull(argv[8], 0, 10) : 0);
	if (argc >= 6 && strcmp(argv[1], "long") == 0)
		return gen_long(argv[2], argv[3], atoll(argv[4]), strtoull(argv[5], 0, 10), argc > 6 ? atof(argv[6]) : 10000.0, argc > 7 ? argv[7] : 0);
	fprintf(stderr, "usage: mmsynth pair <prefix> <bp> <n_ctg> <seed> [edit_seed] [--old]\n"
					"       mmsynth srp  <prefix> <o1.fq> <o2.fq> <n_pairs> <read_seed>\n"
					"       mmsynth long <ref.fa> <out.fq> <n> <seed> [mean [regions.bed]]\n"
					"       mmsynth ref  <out.fa> <bp> <n_ctg> <seed>\n"
					"       mmsynth sr   <ref.fa> <o1.fq> <o2.fq> <n_pairs> <seed> [frac [read_seed]]\n");
	return 2;
}
