#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include "mm2b_priv.h"
/* reads paired files through the read-ahead reader, frees records from OTHER threads in scrambled order, closes early once */
typedef struct { mm_bseq1_t *a; int n; } job_t;
static void *freer(void *p) { job_t *j = (job_t*)p; int i; for (i = j->n - 1; i >= 0; i -= 2) mm_bseq_free1(&j->a[i], 1); for (i = j->n - 2; i >= 0; i -= 2) mm_bseq_free1(&j->a[i], 1); free(j->a); free(j); return 0; }
int main(int argc, char **argv)
{
	int rep;
	for (rep = 0; rep < 2; ++rep) {
		mm_bseq_file_t *fp[2];
		pthread_t th[64]; int nt = 0, n, k = 0; long tot = 0;
		fp[0] = mm_bseq_open(argv[1]); fp[1] = mm_bseq_open(argv[2]);
		mm_bseq_set_readahead(fp[0], 1); mm_bseq_set_readahead(fp[1], 1);
		for (;;) {
			mm_bseq1_t *a = mm_bseq_read_frag2(2, fp, 3000000, 1, rep, &n);
			job_t *j;
			if (a == 0 || n == 0) break;
			tot += n;
			j = (job_t*)malloc(sizeof(job_t)); j->a = a; j->n = n;
			pthread_create(&th[nt++], 0, freer, j);
			if (rep == 1 && ++k == 3) break; /* close before the end of the files */
		}
		mm_bseq_close(fp[0]); mm_bseq_close(fp[1]);
		while (nt > 0) pthread_join(th[--nt], 0);
		fprintf(stderr, "rep %d: %ld records\n", rep, tot);
	}
	return 0;
}
