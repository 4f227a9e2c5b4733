/* TEST / BENCH HARNESS ONLY (never linked into the product).  Times the reference fork's own mapping step -- worker_pipeline
 * step 1 = kt_for(worker_for) -> mm_map_frag (map.c:590-593, 458-498) -- on all host threads, mini-batch by mini-batch, with
 * parsing (step 0) and formatting (step 2) run but timed separately.  Links against _ref/libmm2ref.so, i.e. the unmodified
 * reference objects; the only code here is option handling.
 *
 *   ref_mapstep [-x preset] [-a] [-t threads] [-K minibatch_bases] [-n max_batches] <ref.fa|ref.mmi> <reads_1.fq> [reads_2.fq]
 * prints one JSON object: {"batches":B,"threads":T,"index_s":..,"reads":[..],"read_s":[..],"map_s":[..],"write_s":[..]} */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>
#include "minimap.h"

extern double realtime(void);
int ref_pipeline_steps(const mm_idx_t *mi, const mm_mapopt_t *opt, int n_threads, int n_fp, const char **fn, int mini_batch_size,
                       int max_batches, double *t, long *n);

int main(int argc, char **argv)
{
	mm_idxopt_t ipt;
	mm_mapopt_t opt;
	int c, n_threads = 3, max_batches = 1 << 20, i, nb;
	long K = -1;
	const char *preset = 0;
	mm_idx_reader_t *rd;
	mm_idx_t *mi;
	double t0, *t;
	long *n;
	FILE *out;
	mm_verbose = 2;
	mm_set_opt(0, &ipt, &opt);
	while ((c = getopt(argc, argv, "x:at:K:n:")) >= 0) {
		if (c == 'x') { preset = optarg; if (mm_set_opt(preset, &ipt, &opt) < 0) { fprintf(stderr, "unknown preset\n"); return 1; } }
		else if (c == 'a') opt.flag |= MM_F_OUT_SAM | MM_F_CIGAR;
		else if (c == 't') n_threads = atoi(optarg);
		else if (c == 'K') { char *e; double x = strtod(optarg, &e); if (*e == 'M' || *e == 'm') x *= 1e6; else if (*e == 'G' || *e == 'g') x *= 1e9; else if (*e == 'K' || *e == 'k') x *= 1e3; K = (long)(x + .499); }
		else if (c == 'n') max_batches = atoi(optarg);
	}
	if (K > 0) opt.mini_batch_size = (int)K;
	if (argc - optind < 2) { fprintf(stderr, "usage: ref_mapstep [-x preset] [-a] [-t n] [-K bases] [-n batches] <ref> <reads_1> [reads_2]\n"); return 2; }
	/* the records go where the reference writes them: stdout; keep a handle for the report, send the records to /dev/null */
	out = fdopen(dup(1), "w");
	if (freopen("/dev/null", "w", stdout) == 0) return 1;
	t0 = realtime();
	rd = mm_idx_reader_open(argv[optind], &ipt, 0);
	if (rd == 0) { fprintf(stderr, "cannot open %s\n", argv[optind]); return 1; }
	mi = mm_idx_reader_read(rd, n_threads);
	mm_idx_reader_close(rd);
	if (mi == 0) return 1;
	if ((opt.flag & MM_F_CIGAR) && (mi->flag & MM_I_NO_SEQ)) return 1;
	mm_mapopt_update(&opt, mi);
	t0 = realtime() - t0;
	if (max_batches > 4096) max_batches = 4096;
	t = (double*)calloc((size_t)max_batches * 3, sizeof(double));
	n = (long*)calloc(max_batches, sizeof(long));
	nb = ref_pipeline_steps(mi, &opt, n_threads, argc - optind - 1, (const char**)&argv[optind + 1], opt.mini_batch_size, max_batches, t, n);
	if (nb < 0) return 1;
	fprintf(out, "{\"batches\":%d,\"threads\":%d,\"index_s\":%.3f,\"mid_occ\":%d", nb, n_threads, t0, opt.mid_occ);
	fprintf(out, ",\"reads\":["); for (i = 0; i < nb; ++i) fprintf(out, "%s%ld", i ? "," : "", n[i]);
	fprintf(out, "],\"read_s\":["); for (i = 0; i < nb; ++i) fprintf(out, "%s%.4f", i ? "," : "", t[3 * i]);
	fprintf(out, "],\"map_s\":["); for (i = 0; i < nb; ++i) fprintf(out, "%s%.4f", i ? "," : "", t[3 * i + 1]);
	fprintf(out, "],\"write_s\":["); for (i = 0; i < nb; ++i) fprintf(out, "%s%.4f", i ? "," : "", t[3 * i + 2]);
	fprintf(out, "]}\n");
	fclose(out);
	mm_idx_destroy(mi);
	return 0;
}
