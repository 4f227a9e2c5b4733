/* mmsynth -- seeded synthetic reference / read generator for the AirLift
 * re-alignment workloads named in BASELINE.json.
 *
 *   mmsynth ref   <out.fa> <genome_bp> <n_contigs> <seed>
 *   mmsynth sr    <ref.fa> <out_1.fq> <out_2.fq> <n_pairs> <seed> [region_frac [read_seed]]
 *   mmsynth long  <ref.fa> <out.fq> <n_reads> <seed> [mean_len]
 *
 * ref : uniform random ACGT, split into n_contigs contigs "chr1..", with
 *       repeat families (2 kb blocks x50 per 2 Mbp, mutated 1-10 %) and ~0.1 %
 *       of bases inside short N runs.
 * sr  : 2x150 bp FR pairs, insert ~N(400,50) clipped to [250,800], 50/50 strand,
 *       1 % substitutions, 0.15 % insertions, 0.05 % deletions, drawn from
 *       "updated regions" (random intervals covering region_frac of the genome,
 *       +- one read length) and 3 % of pairs from retired (absent) sequence.
 * long: ONT-shaped reads, log-normal length (mean mean_len, clipped [1k,100k]),
 *       3 % sub + 3 % ins + 3 % del.
 *
 * Everything is a pure function of the arguments (xorshift64* streams).
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
#include <math.h>

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

static const char BASES[5] = "ACGT";
static inline char comp(char c)
{
	switch (c) { case 'A': return 'T'; case 'C': return 'G'; case 'G': return 'C'; case 'T': return 'A'; default: return 'N'; }
}

/* ------------------------------------------------------------------ ref */
typedef struct { int n; char **name; char **seq; int64_t *len; int64_t total; } ref_t;

static int gen_ref(const char *fn, int64_t G, int n_ctg, uint64_t seed)
{
	rng_t r; rng_seed(&r, seed);
	char *g = (char*)malloc(G + 1);
	if (!g) return 1;
	for (int64_t i = 0; i < G; ++i) g[i] = BASES[rng_next(&r) >> 62];
	/* repeat families: per 2 Mbp one 2 kb family with 50 copies at 1-10 % divergence */
	int64_t n_fam = G / 2000000; if (n_fam < 1) n_fam = 1;
	for (int64_t f = 0; f < n_fam; ++f) {
		int64_t src = rng_below(&r, G - 2000);
		double div = 0.01 + 0.09 * rng_unif(&r);
		char blk[2000]; memcpy(blk, g + src, 2000);
		for (int c = 0; c < 50; ++c) {
			int64_t dst = rng_below(&r, G - 2000);
			for (int j = 0; j < 2000; ++j)
				g[dst + j] = rng_unif(&r) < div ? BASES[rng_next(&r) >> 62] : blk[j];
		}
	}
	/* short high-copy family: 300 bp x (G/30000) copies, 1-10 % divergence (fires mid_occ) */
	{
		int64_t src = rng_below(&r, G - 300), copies = G / 30000;
		char blk[300]; memcpy(blk, g + src, 300);
		for (int64_t c = 0; c < copies; ++c) {
			int64_t dst = rng_below(&r, G - 300);
			double div = 0.01 + 0.09 * rng_unif(&r);
			for (int j = 0; j < 300; ++j)
				g[dst + j] = rng_unif(&r) < div ? BASES[rng_next(&r) >> 62] : blk[j];
		}
	}
	/* N runs: ~0.1 % of bases, runs of 10-200 */
	for (int64_t covered = 0; covered < G / 1000; ) {
		int64_t len = 10 + rng_below(&r, 191), at = rng_below(&r, G - len);
		memset(g + at, 'N', len);
		covered += len;
	}
	FILE *fp = fopen(fn, "w");
	if (!fp) { free(g); return 1; }
	int64_t per = G / n_ctg, off = 0;
	for (int c = 0; c < n_ctg; ++c) {
		int64_t len = c == n_ctg - 1 ? G - off : per;
		fprintf(fp, ">chr%d\n", c + 1);
		for (int64_t i = 0; i < len; i += 60) {
			int64_t l = len - i < 60 ? len - i : 60;
			fwrite(g + off + i, 1, l, fp); fputc('\n', fp);
		}
		off += len;
	}
	fclose(fp); free(g);
	return 0;
}

static ref_t *load_ref(const char *fn)
{
	FILE *fp = fopen(fn, "r");
	if (!fp) return 0;
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

static void revcomp(char *s, int l)
{
	for (int i = 0; i < l >> 1; ++i) { char t = comp(s[i]); s[i] = comp(s[l-1-i]); s[l-1-i] = t; }
	if (l & 1) s[l>>1] = comp(s[l>>1]);
}

typedef struct { int ctg; int64_t st, en; } region_t;

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

static int gen_long(const char *ref_fn, const char *fn, int64_t n, uint64_t seed, double mean_len)
{
	ref_t *R = load_ref(ref_fn);
	if (!R) return 1;
	rng_t r; rng_seed(&r, seed);
	FILE *fp = fopen(fn, "w");
	if (!fp) return 1;
	double sigma = 0.6, mu = log(mean_len) - 0.5 * sigma * sigma;
	int cap = 140000; char *out = (char*)malloc(cap + 1), *src = (char*)malloc(110000), *qual = (char*)malloc(cap + 1);
	memset(qual, '5', cap);
	for (int64_t i = 0; i < n; ++i) {
		int len = (int)exp(mu + sigma * rng_norm(&r));
		if (len < 1000) len = 1000; if (len > 100000) len = 100000;
		int c; do c = (int)rng_below(&r, R->n); while (R->len[c] < 2000);
		if (len > R->len[c]) len = (int)R->len[c];
		int64_t st = rng_below(&r, R->len[c] - len + 1);
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
	if (argc >= 6 && strcmp(argv[1], "ref") == 0)
		return gen_ref(argv[2], atoll(argv[3]), atoi(argv[4]), strtoull(argv[5], 0, 10));
	if (argc >= 7 && strcmp(argv[1], "sr") == 0)
		return gen_sr(argv[2], argv[3], argv[4], atoll(argv[5]), strtoull(argv[6], 0, 10), argc > 7 ? atof(argv[7]) : 0.02, argc > 8 ? strtoull(argv[8], 0, 10) : 0);
	if (argc >= 6 && strcmp(argv[1], "long") == 0)
		return gen_long(argv[2], argv[3], atoll(argv[4]), strtoull(argv[5], 0, 10), argc > 6 ? atof(argv[6]) : 10000.0);
	fprintf(stderr, "usage: mmsynth ref <out.fa> <bp> <n_ctg> <seed> | sr <ref.fa> <o1.fq> <o2.fq> <n_pairs> <seed> [frac [read_seed]] | long <ref.fa> <out.fq> <n> <seed> [mean]\n");
	return 2;
}
