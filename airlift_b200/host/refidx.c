/* refidx.c -- the host face of the index: read the reference FASTA, have the device build the minimizer
 * table (mmg_idx_build), keep names / lengths / the 4-bit sequence for the host-side stages.
 * Reference: index.c (mm_idx_reader_*, mm_idx_gen, mm_idx_str, mm_idx_getseq, mm_idx_cal_max_occ, mm_idx_stat). */
#include <stdio.h>
#include <fcntl.h>
#include <unistd.h>
#include "mm2b_priv.h"

static int g_n_dev = 1, g_dev[16] = {0}, g_lanes = 2;

int mm_b200_set_lanes(int lanes)
{
	if (lanes < 1 || lanes > MM_B200_MAX_LANES) return -1;
	g_lanes = lanes;
	return 0;
}

int mm_b200_set_devices(int n_gpus, const int *dev_ids)
{
	int i;
	if (n_gpus < 1 || n_gpus > 16) return -1;
	g_n_dev = n_gpus;
	for (i = 0; i < n_gpus; ++i) g_dev[i] = dev_ids ? dev_ids[i] : i;
	return 0;
}

static void die_gpu(const char *what)
{
	fprintf(stderr, "[ERROR] %s: %s\n", what, mmg_last_error());
	exit(1);
}

/* build the device index from in-memory sequences and wrap it in an mm_idx_t */
static mm_idx_t *idx_from_seqs(int w, int k, int b, int flag, int n, char **seq, const uint32_t *len, char **name)
{
	mm_idx_t *mi = (mm_idx_t*)calloc(1, sizeof(mm_idx_t));
	struct mm_idx_bucket_s *B = (struct mm_idx_bucket_s*)calloc(1, sizeof(*B));
	uint64_t sum = 0;
	int i, d;
	if (k * 2 < b) b = k * 2;
	if (w < 1) w = 1;
	mi->w = w, mi->k = k, mi->b = b, mi->flag = flag, mi->n_seq = n, mi->B = B;
	mi->seq = (mm_idx_seq_t*)calloc(n, sizeof(mm_idx_seq_t));
	for (i = 0; i < n; ++i) {
		mi->seq[i].name = (name && name[i] && !(flag & MM_I_NO_NAME)) ? strdup(name[i]) : 0;
		mi->seq[i].len = len[i], mi->seq[i].offset = sum;
		sum += len[i];
	}
	B->n_dev = g_n_dev;
	B->lanes = g_lanes;
	for (d = 0; d < g_n_dev; ++d) {
		int l;
		pthread_mutex_init(&B->gpu_token[d], 0);
		B->dev_id[d] = g_dev[d];
		for (l = 0; l < B->lanes; ++l)
			if (mmg_init(g_dev[d], &B->ctx[d * B->lanes + l]) != MMG_OK) die_gpu("cannot initialise the GPU");
	}
	if (mmg_idx_build(B->ctx[0], w, k, !!(flag & MM_I_HPC), n, (const char *const*)seq, len, &B->didx[0]) != MMG_OK) die_gpu("index construction failed");
	for (d = 1; d < g_n_dev; ++d) /* replicate over NVLink; nothing else ever crosses GPUs */
		if (mmg_idx_clone_to(B->ctx[d * B->lanes], B->didx[0], &B->didx[d]) != MMG_OK) die_gpu("index replication failed");
	/* the host keeps the packed sequence: CIGAR post-processing and cs/MD read it (index.c:152-162) */
	mi->S = (uint32_t*)calloc((sum + 7) / 8 + 1, 4);
	if (mmg_idx_copy_S(B->didx[0], mi->S) != MMG_OK) die_gpu("cannot fetch the packed reference");
	return mi;
}

void mm_idx_destroy(mm_idx_t *mi)
{
	uint32_t i;
	int d;
	if (mi == 0) return;
	if (mi->B) {
		for (d = 0; d < mi->B->n_dev; ++d) mmg_idx_free(mi->B->didx[d]);
		for (d = 0; d < mi->B->n_dev * mi->B->lanes; ++d) mmg_destroy(mi->B->ctx[d]);
		free(mi->B);
	}
	for (i = 0; i < mi->n_seq; ++i) free(mi->seq[i].name);
	free(mi->seq); free(mi->S); free(mi);
}

mm_idx_t *mm_idx_str(int w, int k, int is_hpc, int bucket_bits, int n, const char **seq, const char **name)
{ /* index.c:385-432 */
	uint32_t *len;
	mm_idx_t *mi;
	int i, flag = 0;
	if (n <= 0) return 0;
	if (is_hpc) flag |= MM_I_HPC;
	if (name == 0) flag |= MM_I_NO_NAME;
	if (bucket_bits < 0) bucket_bits = 14;
	len = (uint32_t*)malloc((size_t)n * 4);
	for (i = 0; i < n; ++i) len[i] = (uint32_t)strlen(seq[i]);
	mi = idx_from_seqs(w, k, bucket_bits, flag, n, (char**)seq, len, (char**)name);
	free(len);
	return mi;
}

int mm_idx_getseq(const mm_idx_t *mi, uint32_t rid, uint32_t st, uint32_t en, uint8_t *seq)
{ /* index.c:152-162 */
	uint64_t i, st1, en1;
	if (rid >= mi->n_seq || st >= mi->seq[rid].len) return -1;
	if (en > mi->seq[rid].len) en = mi->seq[rid].len;
	st1 = mi->seq[rid].offset + st, en1 = mi->seq[rid].offset + en;
	for (i = st1; i < en1; ++i) seq[i - st1] = mm_seq4_get(mi->S, i);
	return (int)(en - st);
}

int32_t mm_idx_cal_max_occ(const mm_idx_t *mi, float f)
{
	int32_t t = 0;
	if (mmg_idx_cal_max_occ(mi->B->didx[0], f, &t) != MMG_OK) die_gpu("mm_idx_cal_max_occ");
	return t;
}

void mm_idx_stat(const mm_idx_t *mi)
{ /* index.c:100-122 */
	uint64_t len = 0;
	uint32_t i;
	const int64_t n = mmg_idx_n_keys(mi->B->didx[0]), sum = mmg_idx_n_minimizers(mi->B->didx[0]);
	extern int64_t mmg_idx_n_singletons(const mmg_idx_t *idx);
	const int64_t n1 = mmg_idx_n_singletons(mi->B->didx[0]);
	fprintf(stderr, "[M::%s] kmer size: %d; skip: %d; is_hpc: %d; #seq: %d\n", __func__, mi->k, mi->w, mi->flag & MM_I_HPC, mi->n_seq);
	for (i = 0; i < mi->n_seq; ++i) len += mi->seq[i].len;
	fprintf(stderr, "[M::%s::%.3f*%.2f] distinct minimizers: %d (%.2f%% are singletons); average occurrences: %.3lf; average spacing: %.3lf\n",
			__func__, realtime() - mm_realtime0, cputime() / (realtime() - mm_realtime0), (int)n, 100.0 * n1 / n, (double)sum / n, (double)len / sum);
}

/* ---- reader (index.c:533-602).  Only FASTA/FASTQ input builds an index here; a prebuilt .mmi is rejected. */

int64_t mm_idx_is_idx(const char *fn)
{
	int fd, is_idx = 0;
	int64_t off_end;
	char magic[4];
	if (strcmp(fn, "-") == 0) return 0;
	fd = open(fn, O_RDONLY);
	if (fd < 0) return -1;
	if ((off_end = lseek(fd, 0, SEEK_END)) >= 4) {
		lseek(fd, 0, SEEK_SET);
		if (read(fd, magic, 4) == 4 && strncmp(magic, MM_IDX_MAGIC, 4) == 0) is_idx = 1;
	}
	close(fd);
	return is_idx ? off_end : 0;
}

mm_idx_reader_t *mm_idx_reader_open(const char *fn, const mm_idxopt_t *opt, const char *fn_out)
{
	const int64_t is_idx = mm_idx_is_idx(fn);
	mm_idx_reader_t *r;
	if (is_idx < 0) return 0;
	if (is_idx > 0) {
		fprintf(stderr, "[ERROR] '%s' is a prebuilt .mmi index; this build constructs its index on the GPU from FASTA (seconds), please pass the FASTA\n", fn);
		return 0;
	}
	if (fn_out) { fprintf(stderr, "[ERROR] -d (index dump) is not supported by this build\n"); return 0; }
	r = (mm_idx_reader_t*)calloc(1, sizeof(mm_idx_reader_t));
	if (opt) r->opt = *opt;
	else mm_idxopt_init(&r->opt);
	r->fp.seq = mm_bseq_open(fn);
	if (r->fp.seq == 0) { free(r); return 0; }
	return r;
}

void mm_idx_reader_close(mm_idx_reader_t *r)
{
	mm_bseq_close(r->fp.seq);
	free(r);
}

int mm_idx_reader_eof(const mm_idx_reader_t *r) { return mm_bseq_eof(r->fp.seq); }

mm_idx_t *mm_idx_reader_read(mm_idx_reader_t *r, int n_threads)
{ /* one index part: sequences are taken in mini-batches until batch_size bases are exceeded (index.c:284-331,353-372) */
	mm_bseq_file_t *fp = r->fp.seq;
	const uint64_t batch_size = r->opt.batch_size;
	const int mini = (uint64_t)r->opt.mini_batch_size < batch_size ? r->opt.mini_batch_size : (int)batch_size;
	char **seq = 0, **name = 0;
	uint32_t *len = 0;
	int n = 0, m = 0, i;
	uint64_t sum_len = 0;
	mm_idx_t *mi;
	if (fp == 0 || mm_bseq_eof(fp)) return 0;
	while (sum_len <= batch_size) {
		int n_seq;
		mm_bseq1_t *s = mm_bseq_read3(fp, mini, 0, 0, 0, &n_seq);
		if (s == 0) break;
		for (i = 0; i < n_seq; ++i) {
			if (n == m) {
				m = m ? m << 1 : 64;
				seq = (char**)realloc(seq, (size_t)m * sizeof(char*)); name = (char**)realloc(name, (size_t)m * sizeof(char*));
				len = (uint32_t*)realloc(len, (size_t)m * 4);
			}
			if (s[i].l_seq == 0 && mm_verbose >= 2) fprintf(stderr, "[WARNING] the length database sequence '%s' is 0\n", s[i].name);
			seq[n] = s[i].seq, name[n] = s[i].name, len[n] = (uint32_t)s[i].l_seq;
			sum_len += (uint64_t)s[i].l_seq;
			++n;
		}
		free(s);
	}
	if (n == 0) return 0;
	mi = idx_from_seqs(r->opt.w, r->opt.k, r->opt.bucket_bits, r->opt.flag, n, seq, len, name);
	for (i = 0; i < n; ++i) { free(seq[i]); free(name[i]); }
	free(seq); free(name); free(len);
	mi->index = r->n_parts++;
	return mi;
}

/* ---- index replication across processes (one process per GPU): rank 0 builds, the others receive the flat
 * device image through a collective the caller runs (bench.py: NCCL broadcast over NVLink) */

int mm_b200_idx_image(const mm_idx_t *mi, mmg_idx_image_t *img) { return mmg_idx_export(mi->B->didx[0], img); }

/* names/lengths come from the FASTA (cheap); the minimizer table, positions and packed sequence are allocated
 * uninitialised on the GPU with the shape given by `shape`; their device pointers are returned in `ptrs` */
mm_idx_t *mm_b200_idx_alloc(const char *fn, const mm_idxopt_t *opt, const mmg_idx_image_t *shape, mmg_idx_image_t *ptrs)
{
	mm_bseq_file_t *fp = mm_bseq_open(fn);
	mm_idx_t *mi;
	struct mm_idx_bucket_s *B;
	int n_seq, i, d = 0, b = opt->bucket_bits;
	uint64_t sum = 0;
	if (fp == 0) return 0;
	mi = (mm_idx_t*)calloc(1, sizeof(mm_idx_t));
	B = (struct mm_idx_bucket_s*)calloc(1, sizeof(*B));
	if (opt->k * 2 < b) b = opt->k * 2;
	mi->w = opt->w < 1 ? 1 : opt->w, mi->k = opt->k, mi->b = b, mi->flag = opt->flag, mi->B = B;
	mi->seq = (mm_idx_seq_t*)calloc(shape->n_seq > 0 ? shape->n_seq : 1, sizeof(mm_idx_seq_t));
	while (d < shape->n_seq) {
		mm_bseq1_t *s = mm_bseq_read3(fp, 1 << 28, 0, 0, 0, &n_seq);
		if (s == 0) break;
		for (i = 0; i < n_seq && d < shape->n_seq; ++i, ++d) {
			mi->seq[d].name = s[i].name, mi->seq[d].len = (uint32_t)s[i].l_seq, mi->seq[d].offset = sum;
			sum += (uint64_t)s[i].l_seq;
			free(s[i].seq);
		}
		for (; i < n_seq; ++i) { free(s[i].seq); free(s[i].name); }
		free(s);
	}
	mm_bseq_close(fp);
	mi->n_seq = d;
	if (d != shape->n_seq || sum != shape->total_len) {
		fprintf(stderr, "[ERROR] '%s' does not match the index being received (%d sequences, %lu bases expected)\n", fn, shape->n_seq, (unsigned long)shape->total_len);
		exit(1);
	}
	B->n_dev = 1, B->lanes = g_lanes, B->dev_id[0] = g_dev[0];
	pthread_mutex_init(&B->gpu_token[0], 0);
	for (i = 0; i < B->lanes; ++i)
		if (mmg_init(g_dev[0], &B->ctx[i]) != MMG_OK) die_gpu("cannot initialise the GPU");
	*ptrs = *shape;
	if (mmg_idx_alloc_like(B->ctx[0], ptrs, &B->didx[0]) != MMG_OK) die_gpu("cannot allocate the index replica");
	return mi;
}

int mm_b200_idx_finalize(mm_idx_t *mi)
{
	uint64_t sum = 0;
	uint32_t i;
	for (i = 0; i < mi->n_seq; ++i) sum += mi->seq[i].len;
	if (mmg_idx_finalize(mi->B->didx[0]) != MMG_OK) die_gpu("index replica");
	mi->S = (uint32_t*)calloc((sum + 7) / 8 + 1, 4);
	if (mmg_idx_copy_S(mi->B->didx[0], mi->S) != MMG_OK) die_gpu("cannot fetch the packed reference");
	return 0;
}
