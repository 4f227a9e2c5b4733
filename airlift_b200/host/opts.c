/* opts.c -- option defaults, presets and validation.  The VALUES are the contract
 * (reference: options.c:5-192); presets are table-driven here. */
#include <stdio.h>
#include "mm2b_priv.h"

void mm_idxopt_init(mm_idxopt_t *io)
{
	memset(io, 0, sizeof(*io));
	io->k = 15, io->w = 10, io->flag = 0, io->bucket_bits = 14;
	io->mini_batch_size = 50000000;
	io->batch_size = 4000000000ULL;
}

void mm_mapopt_init(mm_mapopt_t *mo)
{
	memset(mo, 0, sizeof(*mo));
	mo->seed = 11;
	mo->mid_occ_frac = 2e-4f;
	mo->sdust_thres = 0;
	mo->min_cnt = 3, mo->min_chain_score = 40;
	mo->bw = 500, mo->max_gap = 5000, mo->max_gap_ref = -1;
	mo->max_chain_skip = 25, mo->max_chain_iter = 5000;
	mo->mask_level = 0.5f, mo->pri_ratio = 0.8f, mo->best_n = 5;
	mo->max_join_long = 20000, mo->max_join_short = 2000;
	mo->min_join_flank_sc = 1000, mo->min_join_flank_ratio = 0.5f;
	mo->a = 2, mo->b = 4, mo->q = 4, mo->e = 2, mo->q2 = 24, mo->e2 = 1;
	mo->sc_ambi = 1;
	mo->zdrop = 400, mo->zdrop_inv = 200;
	mo->end_bonus = -1;
	mo->min_dp_max = mo->min_chain_score * mo->a;
	mo->min_ksw_len = 200;
	mo->anchor_ext_len = 20, mo->anchor_ext_shift = 6;
	mo->max_clip_ratio = 1.0f;
	mo->mini_batch_size = 500000000;
	mo->pe_ori = 0, mo->pe_bonus = 33;
}

static void set_scoring(mm_mapopt_t *mo, int a, int b, int q, int q2, int e, int e2)
{
	mo->a = a, mo->b = b, mo->q = q, mo->q2 = q2, mo->e = e, mo->e2 = e2;
}

static void set_asm(mm_idxopt_t *io, mm_mapopt_t *mo, int w, int b, int q, int q2, int e)
{ /* asm5/asm10/asm20 differ only in w and the penalties (options.c:87-104) */
	io->flag = 0, io->k = 19, io->w = w;
	set_scoring(mo, 1, b, q, q2, e, 1);
	mo->zdrop = mo->zdrop_inv = 200;
	mo->min_mid_occ = 100, mo->min_dp_max = 200, mo->best_n = 50;
}

int mm_set_opt(const char *preset, mm_idxopt_t *io, mm_mapopt_t *mo)
{
	if (preset == 0) { mm_idxopt_init(io); mm_mapopt_init(mo); return 0; }
	if (!strcmp(preset, "ava-ont") || !strcmp(preset, "ava-pb")) {
		const int pb = !strcmp(preset, "ava-pb");
		if (pb) io->flag |= MM_I_HPC, io->k = 19, io->w = 5;
		else io->flag = 0, io->k = 15, io->w = 5;
		mo->flag |= MM_F_ALL_CHAINS | MM_F_NO_DIAG | MM_F_NO_DUAL | MM_F_NO_LJOIN;
		mo->min_chain_score = 100, mo->pri_ratio = 0.0f, mo->max_gap = 10000, mo->max_chain_skip = 25;
		if (!pb) mo->bw = 2000;
	} else if (!strcmp(preset, "map10k") || !strcmp(preset, "map-pb")) {
		io->flag |= MM_I_HPC, io->k = 19;
	} else if (!strcmp(preset, "map-ont")) {
		io->flag = 0, io->k = 15;
	} else if (!strcmp(preset, "asm5")) {
		set_asm(io, mo, 19, 19, 39, 81, 3);
	} else if (!strcmp(preset, "asm10")) {
		set_asm(io, mo, 19, 9, 16, 41, 2);
	} else if (!strcmp(preset, "asm20")) {
		set_asm(io, mo, 10, 4, 6, 26, 2);
	} else if (!strcmp(preset, "short") || !strcmp(preset, "sr")) {
		io->flag = 0, io->k = 21, io->w = 11;
		mo->flag |= MM_F_SR | MM_F_FRAG_MODE | MM_F_NO_PRINT_2ND | MM_F_2_IO_THREADS | MM_F_HEAP_SORT;
		mo->pe_ori = 0<<1|1; /* FR */
		set_scoring(mo, 2, 8, 12, 24, 2, 1);
		mo->zdrop = mo->zdrop_inv = 100;
		mo->end_bonus = 10;
		mo->max_frag_len = 800, mo->max_gap = 100, mo->bw = 100;
		mo->pri_ratio = 0.5f;
		mo->min_cnt = 2, mo->min_chain_score = 25, mo->min_dp_max = 40;
		mo->best_n = 20;
		mo->mid_occ = 1000, mo->max_occ = 5000;
		mo->mini_batch_size = 50000000;
	} else if (!strncmp(preset, "splice", 6) || !strcmp(preset, "cdna")) {
		io->flag = 0, io->k = 15, io->w = 5;
		mo->flag |= MM_F_SPLICE | MM_F_SPLICE_FOR | MM_F_SPLICE_REV | MM_F_SPLICE_FLANK;
		mo->max_gap = 2000, mo->max_gap_ref = mo->bw = 200000;
		set_scoring(mo, 1, 2, 2, 32, 1, 0);
		mo->noncan = 9, mo->junc_bonus = 9;
		mo->zdrop = 200, mo->zdrop_inv = 100;
		if (!strcmp(preset, "splice:hq")) mo->junc_bonus = 5, mo->b = 4, mo->q = 6, mo->q2 = 24;
	} else return -1;
	return 0;
}

void mm_mapopt_update(mm_mapopt_t *opt, const mm_idx_t *mi)
{ /* options.c:51-62 */
	if ((opt->flag & MM_F_SPLICE_FOR) || (opt->flag & MM_F_SPLICE_REV)) opt->flag |= MM_F_SPLICE;
	if (opt->mid_occ <= 0) opt->mid_occ = mm_idx_cal_max_occ(mi, opt->mid_occ_frac);
	if (opt->mid_occ < opt->min_mid_occ) opt->mid_occ = opt->min_mid_occ;
	if (mm_verbose >= 3)
		fprintf(stderr, "[M::%s::%.3f*%.2f] mid_occ = %d\n", __func__, realtime() - mm_realtime0, cputime() / (realtime() - mm_realtime0), opt->mid_occ);
}

void mm_mapopt_max_intron_len(mm_mapopt_t *opt, int max_intron_len)
{
	if ((opt->flag & MM_F_SPLICE) && max_intron_len > 0) opt->max_gap_ref = opt->bw = max_intron_len;
}

static int opt_fail(int code, const char *msg)
{
	if (mm_verbose >= 1) fprintf(stderr, "[ERROR]\033[1;31m %s\033[0m\n", msg);
	return code;
}

int mm_check_opt(const mm_idxopt_t *io, const mm_mapopt_t *mo)
{ /* same checks, order, messages and return codes as options.c:137-192 */
	if (mo->split_prefix && (mo->flag & (MM_F_OUT_CS|MM_F_OUT_MD))) return opt_fail(-6, "--cs or --MD doesn't work with --split-prefix");
	if (io->k <= 0 || io->w <= 0) return opt_fail(-5, "-k and -w must be positive");
	if (mo->best_n < 0) return opt_fail(-4, "-N must be no less than 0");
	if (mo->best_n == 0 && mm_verbose >= 2)
		fprintf(stderr, "[WARNING]\033[1;31m '-N 0' reduces mapping accuracy. Please use '--secondary=no' instead.\033[0m\n");
	if (mo->pri_ratio < 0.0f || mo->pri_ratio > 1.0f) return opt_fail(-4, "-p must be within 0 and 1 (including 0 and 1)");
	if ((mo->flag & MM_F_FOR_ONLY) && (mo->flag & MM_F_REV_ONLY)) return opt_fail(-3, "--for-only and --rev-only can't be applied at the same time");
	if (mo->e <= 0 || mo->q <= 0) return opt_fail(-1, "-O and -E must be positive");
	if ((mo->q != mo->q2 || mo->e != mo->e2) && !(mo->e > mo->e2 && mo->q + mo->e < mo->q2 + mo->e2))
		return opt_fail(-2, "dual gap penalties violating E1>E2 and O1+E1<O2+E2");
	if ((mo->q + mo->e) + (mo->q2 + mo->e2) > 127) return opt_fail(-1, "scoring system violating ({-O}+{-E})+({-O2}+{-E2}) <= 127");
	if (mo->zdrop < mo->zdrop_inv) return opt_fail(-5, "Z-drop should not be less than inversion-Z-drop");
	if ((mo->flag & MM_F_NO_PRINT_2ND) && (mo->flag & MM_F_ALL_CHAINS)) return opt_fail(-5, "-X/-P and --secondary=no can't be applied at the same time");
	return 0;
}

void mm_mapopt_to_dev(const mm_mapopt_t *o, mmg_mapopt_t *d)
{
	memset(d, 0, sizeof(*d));
	d->flag = o->flag, d->mid_occ = o->mid_occ, d->max_occ = o->max_occ;
	d->bw = o->bw, d->max_gap = o->max_gap, d->max_gap_ref = o->max_gap_ref, d->max_frag_len = o->max_frag_len;
	d->max_chain_skip = o->max_chain_skip, d->max_chain_iter = o->max_chain_iter, d->min_cnt = o->min_cnt, d->min_chain_score = o->min_chain_score;
	d->pe_ori = o->pe_ori;
	d->a = o->a, d->b = o->b, d->q = o->q, d->e = o->e, d->q2 = o->q2, d->e2 = o->e2, d->sc_ambi = o->sc_ambi;
}
