/* llsw.c -- local-alignment score with end coordinates, as used by the inversion test of
 * mm_test_zdrop (align.c:72-87) and by mm_align1_inv (align.c:812-820).
 *
 * The reference computes it with a striped (Farrar) int16 Smith-Waterman: ksw_ll_qinit +
 * ksw_ll_i16, ksw2_ll_sse.c:32-147.  Striping is observable: the vertical gap state E is updated
 * from H before the lazy-F correction, the query is zero-padded to a multiple of 8, ties on the
 * best row resolve to the LAST row, and the query end is the last maximal cell in striped memory
 * order.  So the eight lanes are walked here exactly as the vector code walks them. */
#include "mm2b_priv.h"

#define LANES 8

MM_FN static inline int16_t sat_add16(int16_t a, int16_t b)
{
	int32_t s = (int32_t)a + b;
	return (int16_t)(s > 32767 ? 32767 : s < -32768 ? -32768 : s);
}
MM_FN static inline int16_t sat_subu16(int16_t a, int16_t b) /* unsigned saturating a-b on the bit patterns */
{
	uint16_t ua = (uint16_t)a, ub = (uint16_t)b;
	return (int16_t)(ua > ub ? ua - ub : 0);
}
MM_FN static inline int16_t max16(int16_t a, int16_t b) { return a > b ? a : b; }

MM_FN int mm_ll_i16(int qlen, const uint8_t *query, int tlen, const uint8_t *target, int m, const int8_t *mat, int gapo, int gape, int *qe, int *te)
{
	const int slen = (qlen + LANES - 1) / LANES;
	const int16_t gapoe = (int16_t)(gapo + gape), ge = (int16_t)gape;
	int16_t *prof, *H0, *H1, *E, *Hmax, *tmp;
	int a, i, j, k, l, gmax = 0;
	*qe = *te = -1;
	if (qlen <= 0) return 0;
	prof = (int16_t*)malloc((size_t)slen * LANES * (m + 4) * sizeof(int16_t));
	H0 = prof + (size_t)slen * LANES * m, H1 = H0 + slen * LANES, E = H1 + slen * LANES, Hmax = E + slen * LANES;
	for (a = 0; a < m; ++a) /* striped query profile: vector j, lane l holds query position j + l*slen (ksw2_ll_sse.c:70-78) */
		for (j = 0; j < slen; ++j)
			for (l = 0; l < LANES; ++l) {
				const int p = j + l * slen;
				prof[((size_t)a * slen + j) * LANES + l] = p >= qlen ? 0 : mat[a * m + query[p]];
			}
	memset(H0, 0, (size_t)slen * LANES * 2); memset(E, 0, (size_t)slen * LANES * 2); memset(Hmax, 0, (size_t)slen * LANES * 2);
	for (i = 0; i < tlen; ++i) {
		int16_t h[LANES], f[LANES], mx[LANES], e;
		const int16_t *S = prof + (size_t)target[i] * slen * LANES;
		int imax = 0, done = 0;
		for (l = 0; l < LANES; ++l) f[l] = 0, mx[l] = 0;
		h[0] = 0; /* H0[slen-1] shifted up by one lane */
		for (l = 1; l < LANES; ++l) h[l] = H0[(size_t)(slen - 1) * LANES + l - 1];
		for (j = 0; j < slen; ++j) {
			for (l = 0; l < LANES; ++l) {
				int16_t hh = sat_add16(h[l], S[(size_t)j * LANES + l]);
				e = E[(size_t)j * LANES + l];
				hh = max16(hh, e); hh = max16(hh, f[l]);
				mx[l] = max16(mx[l], hh);
				H1[(size_t)j * LANES + l] = hh;
				hh = sat_subu16(hh, gapoe);
				e = sat_subu16(e, ge); e = max16(e, hh);
				E[(size_t)j * LANES + l] = e;
				f[l] = sat_subu16(f[l], ge); f[l] = max16(f[l], hh);
				h[l] = H0[(size_t)j * LANES + l];
			}
		}
		for (k = 0; k < LANES && !done; ++k) { /* lazy F across lane boundaries (ksw2_ll_sse.c:125-135) */
			for (l = LANES - 1; l > 0; --l) f[l] = f[l - 1];
			f[0] = 0;
			for (j = 0; j < slen; ++j) {
				int any = 0;
				for (l = 0; l < LANES; ++l) {
					int16_t hh = max16(H1[(size_t)j * LANES + l], f[l]);
					H1[(size_t)j * LANES + l] = hh;
					hh = sat_subu16(hh, gapoe);
					f[l] = sat_subu16(f[l], ge);
					if (f[l] > hh) any = 1;
				}
				if (!any) { done = 1; break; }
			}
		}
		for (l = 0; l < LANES; ++l) imax = imax > mx[l] ? imax : mx[l];
		if (imax >= gmax) {
			gmax = imax, *te = i;
			memcpy(Hmax, H1, (size_t)slen * LANES * 2);
		}
		tmp = H1, H1 = H0, H0 = tmp;
	}
	for (i = 0; i < slen * LANES; ++i)
		if ((int)(uint16_t)Hmax[i] == gmax) *qe = i / LANES + i % LANES * slen;
	free(prof);
	return gmax;
}
