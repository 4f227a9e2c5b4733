/* seqio.c -- gz FASTA/FASTQ reader and mini-batch assembly.
 * Observable behaviour follows the reference's kseq.h:192-233 (record grammar) and bseq.c:58-162
 * (U->T, batch sizing, mate grouping); SURVEY.md App. C lists the rules.  Implementation is a small
 * pull parser over a 256 KiB inflate window. */
#include <zlib.h>
#include <stdio.h>
#include <ctype.h>
#include "mm2b_priv.h"

#define IO_BUF (256 * 1024)
#define CHECK_PAIR_THRES 1000000

typedef struct {
	gzFile fp;
	unsigned char *buf;
	int beg, end, eof;
} gzstream_t;

typedef struct { mm_str_t name, comment, seq, qual; int last_char; } record_t;

struct mm_bseq_file_s {
	gzstream_t st;
	record_t rec;
	mm_bseq1_t pending; /* a record read ahead while completing a pair (bseq.c:100-110) */
};

static inline int st_fill(gzstream_t *s)
{
	if (s->eof) return 0;
	s->beg = 0;
	s->end = gzread(s->fp, s->buf, IO_BUF);
	if (s->end < IO_BUF) s->eof = 1;
	if (s->end < 0) s->end = 0;
	return s->end;
}

static inline int st_getc(gzstream_t *s)
{
	if (s->beg >= s->end && st_fill(s) == 0) return -1;
	return s->buf[s->beg++];
}

static inline void str_reserve(mm_str_t *s, size_t need)
{
	if (s->m < need) { s->m = mm_roundup32((uint32_t)need); s->s = (char*)realloc(s->s, s->m); }
}

/* read up to a delimiter (is_line: '\n', else any whitespace); the delimiter is consumed and returned in *dret.
 * Returns the string length or -1 at end of file.  A single trailing '\r' is dropped from line reads. */
static int st_until(gzstream_t *s, int is_line, mm_str_t *str, int *dret, int append)
{
	if (dret) *dret = 0;
	if (!append) str->l = 0;
	if (s->beg >= s->end && s->eof) return -1;
	for (;;) {
		int i;
		if (s->beg >= s->end && st_fill(s) == 0) break;
		if (is_line) { unsigned char *p = (unsigned char*)memchr(s->buf + s->beg, '\n', s->end - s->beg); i = p ? (int)(p - s->buf) : s->end; }
		else for (i = s->beg; i < s->end; ++i) if (isspace(s->buf[i])) break;
		str_reserve(str, (size_t)str->l + (i - s->beg) + 1);
		memcpy(str->s + str->l, s->buf + s->beg, i - s->beg);
		str->l += i - s->beg;
		s->beg = i + 1;
		if (i < s->end) { if (dret) *dret = s->buf[i]; break; }
	}
	if (str->s == 0) { str->m = 1; str->s = (char*)calloc(1, 1); }
	else if (is_line && str->l > 1 && str->s[str->l - 1] == '\r') --str->l;
	str->s[str->l] = 0;
	return (int)str->l;
}

/* >=0 sequence length; -1 end of file; -2 truncated/mismatched quality */
static int read_record(gzstream_t *st, record_t *r)
{
	int c;
	if (r->last_char == 0) {
		while ((c = st_getc(st)) != -1 && c != '>' && c != '@') {}
		if (c == -1) return -1;
		r->last_char = c;
	}
	r->comment.l = r->seq.l = r->qual.l = 0;
	if (st_until(st, 0, &r->name, &c, 0) < 0) return -1;
	if (c != '\n') st_until(st, 1, &r->comment, 0, 0);
	if (r->seq.s == 0) { r->seq.m = 256; r->seq.s = (char*)malloc(r->seq.m); }
	while ((c = st_getc(st)) != -1 && c != '>' && c != '+' && c != '@') {
		if (c == '\n') continue;
		str_reserve(&r->seq, (size_t)r->seq.l + 2);
		r->seq.s[r->seq.l++] = (char)c;
		st_until(st, 1, &r->seq, 0, 1);
	}
	if (c == '>' || c == '@') r->last_char = c;
	str_reserve(&r->seq, (size_t)r->seq.l + 2);
	r->seq.s[r->seq.l] = 0;
	if (c != '+') return (int)r->seq.l;
	str_reserve(&r->qual, r->seq.m);
	while ((c = st_getc(st)) != -1 && c != '\n') {}
	if (c == -1) return -2;
	while (st_until(st, 1, &r->qual, 0, 1) >= 0 && r->qual.l < r->seq.l) {}
	r->last_char = 0;
	if (r->seq.l != r->qual.l) return -2;
	return (int)r->seq.l;
}

mm_bseq_file_t *mm_bseq_open(const char *fn)
{
	gzFile f = fn && strcmp(fn, "-") ? gzopen(fn, "r") : gzdopen(0, "r");
	mm_bseq_file_t *fp;
	if (f == 0) return 0;
	gzbuffer(f, 1 << 20);
	fp = (mm_bseq_file_t*)calloc(1, sizeof(*fp));
	fp->st.fp = f;
	fp->st.buf = (unsigned char*)malloc(IO_BUF);
	return fp;
}

void mm_bseq_close(mm_bseq_file_t *fp)
{
	if (fp == 0) return;
	free(fp->rec.name.s); free(fp->rec.comment.s); free(fp->rec.seq.s); free(fp->rec.qual.s);
	free(fp->st.buf);
	gzclose(fp->st.fp);
	free(fp);
}

int mm_bseq_eof(mm_bseq_file_t *fp)
{
	return fp->st.eof && fp->st.beg >= fp->st.end && fp->pending.seq == 0;
}

static char *dup_str(const mm_str_t *s)
{
	char *t = (char*)malloc(s->l + 1);
	memcpy(t, s->s, s->l); t[s->l] = 0;
	return t;
}

static void record_to_bseq(const record_t *r, mm_bseq1_t *s, int with_qual, int with_comment)
{
	int i;
	if (r->name.l == 0) fprintf(stderr, "[WARNING]\033[1;31m empty sequence name in the input.\033[0m\n");
	s->name = dup_str(&r->name);
	s->seq = dup_str(&r->seq);
	for (i = 0; i < (int)r->seq.l; ++i) /* U -> T, u -> t (bseq.c:72-74) */
		if (s->seq[i] == 'u' || s->seq[i] == 'U') --s->seq[i];
	s->qual = with_qual && r->qual.l ? dup_str(&r->qual) : 0;
	s->comment = with_comment && r->comment.l ? dup_str(&r->comment) : 0;
	s->l_seq = (int)r->seq.l;
	s->rid = 0;
}

typedef struct { int n, m; mm_bseq1_t *a; } bseq_v;

static inline mm_bseq1_t *vec_next(bseq_v *v)
{
	if (v->n == v->m) { v->m = v->m ? v->m << 1 : 256; v->a = (mm_bseq1_t*)realloc(v->a, (size_t)v->m * sizeof(mm_bseq1_t)); }
	return &v->a[v->n++];
}

int mm_qname_len(const char *s)
{ /* a trailing "/<digit>" is not part of the fragment name (bseq.h:31-36) */
	int l = (int)strlen(s);
	return l >= 3 && s[l-1] >= '0' && s[l-1] <= '9' && s[l-2] == '/' ? l - 2 : l;
}

int mm_qname_same(const char *s1, const char *s2)
{
	int l1 = mm_qname_len(s1), l2 = mm_qname_len(s2);
	return l1 == l2 && strncmp(s1, s2, l1) == 0;
}

void mm_revcomp_bseq(mm_bseq1_t *s)
{ /* bseq.h:46-58 */
	int i, l = s->l_seq;
	for (i = 0; i < l >> 1; ++i) {
		int t = (unsigned char)s->seq[l - i - 1];
		s->seq[l - i - 1] = seq_comp_table[(unsigned char)s->seq[i]];
		s->seq[i] = seq_comp_table[t];
	}
	if (l & 1) s->seq[l >> 1] = seq_comp_table[(unsigned char)s->seq[l >> 1]];
	if (s->qual)
		for (i = 0; i < l >> 1; ++i) { char t = s->qual[l - i - 1]; s->qual[l - i - 1] = s->qual[i], s->qual[i] = t; }
}

mm_bseq1_t *mm_bseq_read3(mm_bseq_file_t *fp, int chunk_size, int with_qual, int with_comment, int frag_mode, int *n_)
{ /* bseq.c:80-117 */
	int64_t size = 0;
	int ret;
	bseq_v a = {0, 0, 0};
	*n_ = 0;
	if (fp->pending.seq) {
		*vec_next(&a) = fp->pending;
		size = fp->pending.l_seq;
		memset(&fp->pending, 0, sizeof(mm_bseq1_t));
	}
	while ((ret = read_record(&fp->st, &fp->rec)) >= 0) {
		mm_bseq1_t *s = vec_next(&a);
		record_to_bseq(&fp->rec, s, with_qual, with_comment);
		size += s->l_seq;
		if (size >= chunk_size) {
			if (frag_mode && a.a[a.n-1].l_seq < CHECK_PAIR_THRES) { /* never split a pair across batches */
				while (read_record(&fp->st, &fp->rec) >= 0) {
					record_to_bseq(&fp->rec, &fp->pending, with_qual, with_comment);
					if (mm_qname_same(fp->pending.name, a.a[a.n-1].name)) {
						*vec_next(&a) = fp->pending;
						memset(&fp->pending, 0, sizeof(mm_bseq1_t));
					} else break;
				}
			}
			break;
		}
	}
	if (ret < -1) fprintf(stderr, "[WARNING]\033[1;31m wrong FASTA/FASTQ record. Continue anyway.\033[0m\n");
	*n_ = a.n;
	return a.a;
}

mm_bseq1_t *mm_bseq_read_frag2(int n_fp, mm_bseq_file_t **fp, int chunk_size, int with_qual, int with_comment, int *n_)
{ /* bseq.c:129-157: zip the files record by record; stop at the shortest */
	int i;
	int64_t size = 0;
	bseq_v a = {0, 0, 0};
	*n_ = 0;
	if (n_fp < 1) return 0;
	for (;;) {
		int n_read = 0;
		for (i = 0; i < n_fp; ++i)
			if (read_record(&fp[i]->st, &fp[i]->rec) >= 0) ++n_read;
		if (n_read < n_fp) {
			if (n_read > 0)
				fprintf(stderr, "[W::%s]\033[1;31m query files have different number of records; extra records skipped.\033[0m\n", __func__);
			break;
		}
		for (i = 0; i < n_fp; ++i) {
			mm_bseq1_t *s = vec_next(&a);
			record_to_bseq(&fp[i]->rec, s, with_qual, with_comment);
			size += s->l_seq;
		}
		if (size >= chunk_size) break;
	}
	*n_ = a.n;
	return a.a;
}
