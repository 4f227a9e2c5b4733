// mmg_post.h -- the post-chaining stages of one shard as data-parallel steps over flat arrays.
//
// What mm_map_frag does between mm_chain_dp and its return (map.c:376-400) -- hits, primary/secondary tree, per-mate split,
// base-level alignment, filters, MAPQ, pairing -- is cut here into steps whose work items are independent:
//
//   per chain     post_hit_record     sorted chain keys -> hit records with coordinates            (mm_gen_regs)
//   per fragment  post_hits_select    primary/secondary tree, secondary selection                   (chain_post)
//   per fragment  post_mates          per-mate hit lists and anchors, per-mate tree                 (mm_seg_gen + mm_set_parent)
//   per read      post_plan           anchor stretch, extension windows, DP jobs of every new hit   (mm_align1, first half)
//   per read      post_size/build     CIGAR stitching, coordinates, clean-up, z-drop cuts           (mm_align1, second half)
//   per read      post_final          filters, final order, tree, selection                         (mm_align_skeleton tail, align_regs)
//   per fragment  post_finish         MAPQ, pairing, mate un-flip                                   (mm_set_mapq, mm_pair, map.c:486-497)
//
// All state lives in arrays sized by prefix sums; a step reads and writes only the slots of its own item.  The kernels of
// mmg_post.cu are one-line wrappers around these functions; tests/emu/ runs the same functions on the CPU.
#pragma once
#include "mmg_aln.h"

struct PostShard {
	HitOpt opt;
	LogTab lt;
	int32_t nf, n_seq, idx_k, pad0;
	// the resident batch
	const int32_t *n_seg, *seg_off, *seq_len;    // per fragment, per fragment, per read
	const uint64_t *q_off;                        // per read
	const uint8_t *flip;                          // per read: mate stored reverse-complemented (map.c:467-469)
	const uint32_t *Q, *S;
	const uint64_t *ref_off;
	const uint32_t *ref_len;
	// chaining output
	const int32_t *nu, *rep;                      // per fragment
	const int64_t *uoff, *voff;                   // per fragment (+1)
	const uint64_t *u;
	mm128 *a;
	const uint32_t *hash;                         // per fragment: salt of map.c:291-293
	// fragment-level hits: slots uoff[f] .. uoff[f] + nu[f]
	const uint64_t *key_in;                       // keys as computed, chain i of fragment f at uoff[f] + nu[f] - 1 - i
	uint64_t *key, *ascnt;                        // keys in hit order and their (first anchor << 32 | count) words
	HitRec *r0;
	int32_t *w; uint64_t *cov; mm128 *big; RsFrame *stack;   // scratch, same slots (stack: uoff[f] / 65 + 2 f)
	int32_t *n0;                                  // per fragment: hits kept by the selection
	// per-read hit lists: slots roff[read] .. roff[read] + cap[read]
	const int64_t *roff;
	int32_t *cap, *n_reg;
	HitRec *r1, *tmp1;
	RegionPlan *pl;
	uint32_t *xsize; const int64_t *xoff;         // words of each hit's alignment record, and where it starts in xw (relative to x_base)
	int64_t x_base;                               // words of xw taken by earlier rounds
	uint64_t *skey; int32_t *sidx; mm128 *sbig; RsFrame *sstack;   // scratch (sstack: roff[read] / 65 + 2 read)
	mm128 *a1;                                    // per-mate anchors: fragment f owns voff[f] .. ; read's share starts at a1_off[read]
	int64_t *a1_off;
	// DP
	DpJob *jobs; unsigned int *n_jobs; unsigned int job_cap;
	DpView dp;
	uint32_t *xw;
	// counters: [0] hits still to be aligned, [1] error flags
	unsigned int *ctr;
};
#define POST_ERR_SLOTS   1u   // more z-drop cuts in one read than spare hit slots
#define POST_ERR_JOBS    2u   // DP job buffer full
#define POST_ERR_LOGTAB  4u   // a logf argument outside the table
// hit slots of a read: every hit kept by the selection may be cut once at a z-drop, plus two
#define POST_SLOTS(n0) (2 * (n0) + 2)

MMG_HD int post_frag_of(const int64_t *off, int n, int64_t g)
{ // last f with off[f] <= g
	int lo = 0, hi = n - 1;
	while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (off[mid] <= g) lo = mid; else hi = mid - 1; }
	return lo;
}

MMG_HD int post_qlen_sum(const PostShard &sh, int f)
{
	int s = 0;
	for (int j = 0; j < sh.n_seg[f]; ++j) s += sh.seq_len[sh.seg_off[f] + j];
	return s;
}

MMG_HD int post_frag_gap(const HitOpt &o, int qlen_sum)
{ // map.c:344-349
	if (o.max_gap_ref > 0) return o.max_gap_ref;
	if (o.max_frag_len > 0) { const int g = o.max_frag_len - qlen_sum; return g < o.max_gap ? o.max_gap : g; }
	return o.max_gap;
}

MMG_HD bool post_frag_mapped(const PostShard &sh, int f, int qlen_sum)
{ // map.c:283-284
	const int ns = sh.n_seg[f];
	return !(qlen_sum == 0 || ns <= 0 || ns > 8 || (sh.opt.max_qlen > 0 && qlen_sum > sh.opt.max_qlen));
}

// key of chain g (global index) written where the stable descending sort leaves klib's order (see hit_order_desc):
// the chains of a fragment are handed to the sort back to front
MMG_HD void post_chain_key(const PostShard &sh, int64_t g, const int64_t *cnt_prefix, uint64_t *key_in, uint64_t *val_in)
{
	const int f = post_frag_of(sh.uoff, sh.nf, g);
	const int64_t i = g - sh.uoff[f], first = cnt_prefix[g] - sh.voff[f];
	const int64_t slot = sh.uoff[f] + (sh.nu[f] - 1 - i);
	key_in[slot] = hit_key(sh.u[g], sh.a[sh.voff[f] + first], sh.hash[f]);
	val_in[slot] = (uint64_t)first << 32 | (uint32_t)(int32_t)sh.u[g];
}

MMG_HD void post_hit_record(const PostShard &sh, int64_t g)
{
	const int f = post_frag_of(sh.uoff, sh.nf, g);
	hit_init(&sh.r0[g], (int32_t)(g - sh.uoff[f]), sh.key[g], sh.ascnt[g], post_qlen_sum(sh, f), sh.a + sh.voff[f]);
}

// fragments of more than 64 chains: equal keys leave klib's radix sort in an order only its replay gives (hit_order_desc)
MMG_HDN inline void post_hits_fix_order(const PostShard &sh, int f, int known_tie = -1 /* 1: the caller found equal keys; 0: none; -1: look */)
{
	const int n = sh.nu[f];
	if (n <= 64) return;
	const int64_t o = sh.uoff[f];
	bool tie = known_tie > 0;
	if (known_tie < 0) for (int i = 1; i < n && !tie; ++i) tie = sh.key[o + i] == sh.key[o + i - 1];
	if (!tie) return;
	// original order: key_in holds the chains back to front
	uint64_t *k = sh.cov + o;
	for (int i = 0; i < n; ++i) k[i] = sh.key_in[o + n - 1 - i];
	hit_order_desc(n, k, sh.w + o, sh.big + o, sh.stack + o / 65 + 2 * (int64_t)f);
	const int qlen = post_qlen_sum(sh, f);
	int64_t first = 0;
	for (int i = 0; i < n; ++i) { sh.big[o + i].y = (uint64_t)first << 32 | (uint32_t)(int32_t)sh.u[o + i]; first += (int32_t)sh.u[o + i]; }
	for (int r = 0; r < n; ++r) { const int i = sh.w[o + r]; hit_init(&sh.r0[o + r], r, k[i], sh.big[o + i].y, qlen, sh.a + sh.voff[f]); }
}

// chain_post (map.c:249-258) without the long join (not part of the short-read presets)
MMG_HDN inline void post_hits_select(const PostShard &sh, int f)
{
	const int ns = sh.n_seg[f], qlen_sum = post_qlen_sum(sh, f);
	const int64_t o = sh.uoff[f];
	int n = sh.nu[f];
	if (!post_frag_mapped(sh, f, qlen_sum)) n = 0;
	if (n > 0) {
		post_hits_fix_order(sh, f);
		if (!(sh.opt.flag & HIT_F_ALL_CHAINS)) {
			hit_set_parent(sh.opt.mask_level, n, sh.r0 + o, sh.opt.a * 2 + sh.opt.b, (sh.opt.flag & HIT_F_HARD_MLEVEL) != 0, sh.w + o, sh.cov + o, nullptr);
			if (ns <= 1) n = hit_select_sub(sh.opt.pri_ratio, sh.idx_k * 2, sh.opt.best_n, n, sh.r0 + o, sh.w + o);
			else {
				int32_t ql[8];
				for (int j = 0; j < ns; ++j) ql[j] = sh.seq_len[sh.seg_off[f] + j];
				n = hit_select_sub_multi(sh.opt.pri_ratio, 0.2f, 0.7f, post_frag_gap(sh.opt, qlen_sum), sh.idx_k * 2, sh.opt.best_n, ns, ql, n, sh.r0 + o, sh.w + o);
			}
		}
	}
	sh.n0[f] = n;
	for (int j = 0; j < ns; ++j) sh.cap[sh.seg_off[f] + j] = n > 0 ? POST_SLOTS(n) : 0;
}

// the same step for a fragment with many chains, run by one warp: the tree is built cooperatively (hit_set_parent_warp), the
// selection stays with one lane (it is a single pass whose reads depend on its own earlier writes, pe.c:13-40)
#define POST_WARP_MIN_CHAINS 32
template <class W>
MMG_HDN inline void post_hits_select_warp(const W &wp, const PostShard &sh, int f, int32_t *fast /* HIT_FAST_WORDS + 2 ints */)
{
	const int ns = sh.n_seg[f], qlen_sum = post_qlen_sum(sh, f);
	const int64_t o = sh.uoff[f];
	const int n_in = post_frag_mapped(sh, f, qlen_sum) ? sh.nu[f] : 0;
	const bool tree = n_in > 0 && !(sh.opt.flag & HIT_F_ALL_CHAINS);
	if (n_in > 64) { // equal keys among the sorted chain keys?  (one lane reading them one after the other held a 10^4-chain fragment for most of its time here)
		bool tie = false;
		for (int i0 = 1; i0 < n_in && !tie; i0 += 32)
			tie = wp.ballot([&](int l) { const int i = i0 + l; return i < n_in && sh.key[o + i] == sh.key[o + i - 1]; }) != 0;
		if (tie) wp.one([&]() { post_hits_fix_order(sh, f, 1); });
	}
	if (tree) {
		if (qlen_sum <= HIT_COVER_BITS) hit_set_parent_warp(wp, sh.opt.mask_level, n_in, sh.r0 + o, (sh.opt.flag & HIT_F_HARD_MLEVEL) != 0, sh.w + o, fast);
		else wp.one([&]() { hit_set_parent(sh.opt.mask_level, n_in, sh.r0 + o, sh.opt.a * 2 + sh.opt.b, (sh.opt.flag & HIT_F_HARD_MLEVEL) != 0, sh.w + o, sh.cov + o, nullptr); });
	}
	int n = n_in;
	if (tree) {
		if (ns <= 1) wp.one([&]() { fast[HIT_FAST_WORDS] = hit_select_sub(sh.opt.pri_ratio, sh.idx_k * 2, sh.opt.best_n, n_in, sh.r0 + o, sh.w + o); });
		else {
			int32_t ql[8];
			for (int j = 0; j < ns; ++j) ql[j] = sh.seq_len[sh.seg_off[f] + j];
			n = hit_select_sub_multi_warp(wp, sh.opt.pri_ratio, 0.2f, 0.7f, post_frag_gap(sh.opt, qlen_sum), sh.idx_k * 2, sh.opt.best_n, ns, ql, n_in, sh.r0 + o, sh.w + o, fast + HIT_FAST_WORDS);
		}
		if (ns <= 1) n = fast[HIT_FAST_WORDS];
	}
	wp.one([&]() {
		sh.n0[f] = n;
		for (int j = 0; j < ns; ++j) sh.cap[sh.seg_off[f] + j] = n > 0 ? POST_SLOTS(n) : 0;
	});
}

// mm_seg_gen + per-mate mm_set_parent (map.c:395-399), or the single-segment hand-over (map.c:390-393)
MMG_HDN inline void post_mates(const PostShard &sh, int f)
{
	const int ns = sh.n_seg[f], off = sh.seg_off[f], n0 = sh.n0[f];
	const int64_t o = sh.uoff[f];
	for (int j = 0; j < ns; ++j) sh.n_reg[off + j] = 0, sh.a1_off[off + j] = sh.voff[f];
	if (n0 == 0) return;
	if (ns == 1) {
		HitRec *r = sh.r1 + sh.roff[off];
		for (int i = 0; i < n0; ++i) r[i] = sh.r0[o + i], sh.pl[sh.roff[off] + i].state = 0;
		sh.n_reg[off] = n0;
		// the read's anchors stay where chaining left them; mm_align_skeleton compacts them first (align.c:873)
		if (sh.opt.flag & HIT_F_CIGAR) hit_squeeze(n0, r, sh.a + sh.voff[f], sh.cov + o);
		return;
	}
	int32_t ql[8], seg_a0[9], n_out[8];
	HitRec *r_out[8];
	for (int j = 0; j < ns; ++j) ql[j] = sh.seq_len[off + j], r_out[j] = sh.r1 + sh.roff[off + j];
	hit_split_mates(sh.hash[f], ns, ql, n0, sh.r0 + o, sh.a + sh.voff[f], sh.a1 + sh.voff[f], seg_a0, r_out, n_out, sh.key + o, sh.cov + o, sh.w + o, sh.big + o,
	                sh.stack + o / 65 + 2 * (int64_t)f);
	for (int j = 0; j < ns; ++j) {
		sh.n_reg[off + j] = n_out[j], sh.a1_off[off + j] = sh.voff[f] + seg_a0[j];
		for (int i = 0; i < n_out[j]; ++i) sh.pl[sh.roff[off + j] + i].state = 0;
		hit_set_parent(sh.opt.mask_level, n_out[j], r_out[j], sh.opt.a * 2 + sh.opt.b, (sh.opt.flag & HIT_F_HARD_MLEVEL) != 0, sh.sidx + sh.roff[off + j], sh.skey + sh.roff[off + j], nullptr);
	}
}

MMG_HD const mm128 *post_read_anchors(const PostShard &sh, int f, int read) { return (sh.n_seg[f] == 1 ? sh.a : sh.a1) + sh.a1_off[read]; }

MMG_HD SeqView post_seq_view(const PostShard &sh, int read)
{
	SeqView v;
	v.Q = sh.Q, v.q_off = sh.q_off[read], v.qlen = sh.seq_len[read], v.S = sh.S, v.ref_off = sh.ref_off, v.ref_len = sh.ref_len;
	return v;
}

struct PostEmit { // DP job slots are handed out by one atomic counter; the order of jobs carries no meaning
	const PostShard *sh;
	MMG_HD int32_t operator()(const DpJob &j) const
	{
#ifdef __CUDA_ARCH__
		const unsigned int id = atomicAdd(sh->n_jobs, 1u);
#else
		const unsigned int id = (*sh->n_jobs)++;
#endif
		if (id >= sh->job_cap) {
#ifdef __CUDA_ARCH__
			atomicOr(sh->ctr + 1, POST_ERR_JOBS);
#else
			sh->ctr[1] |= POST_ERR_JOBS;
#endif
			return -1;
		}
		sh->jobs[id] = j;
		return (int32_t)id;
	}
};

// every hit of the read that has not been planned yet: stretch, windows, jobs.  f: the read's fragment.
MMG_HDN inline void post_plan(const PostShard &sh, int f, int read)
{
	const int n = sh.n_reg[read];
	if (n == 0) return;
	const int64_t ro = sh.roff[read];
	const SeqView v = post_seq_view(sh, read);
	const mm128 *a = post_read_anchors(sh, f, read);
	PostEmit emit = {&sh};
	for (int i = 0; i < n; ++i)
		if (sh.pl[ro + i].state == 0) aln_plan_sr(sh.opt, v, read, sh.r1[ro + i], a, &sh.pl[ro + i], emit);
}

MMG_HDN inline void post_size(const PostShard &sh, int read)
{
	const int64_t ro = sh.roff[read];
	for (int i = 0; i < sh.cap[read]; ++i)
		sh.xsize[ro + i] = (i < sh.n_reg[read] && sh.pl[ro + i].state == 1) ? aln_record_words(sh.pl[ro + i], sh.dp) : 0;
}

// build every planned hit; tail hits of z-drop cuts are slotted in right behind the hit they were cut from (align.c:906,
// mm_insert_reg) and wait for the next round
MMG_HDN inline void post_build(const PostShard &sh, int f, int read)
{
	int n = sh.n_reg[read];
	if (n == 0) return;
	const int64_t ro = sh.roff[read];
	const int cap = sh.cap[read];
	const SeqView v = post_seq_view(sh, read);
	const mm128 *a = post_read_anchors(sh, f, read);
	int n_new = 0;
	for (int i = 0; i < n; ++i) {
		RegionPlan *pl = &sh.pl[ro + i];
		if (pl->state != 1) continue;
		HitRec r2;
		const int cut = aln_build_sr(sh.opt, v, &sh.r1[ro + i], &r2, a, *pl, sh.dp, sh.xw, (uint64_t)(sh.x_base + sh.xoff[ro + i]));
		pl->state = 2;
		if (cut) {
			if (n + n_new >= cap) {
#ifdef __CUDA_ARCH__
				atomicOr(sh.ctr + 1, POST_ERR_SLOTS);
#else
				sh.ctr[1] |= POST_ERR_SLOTS;
#endif
				continue;
			}
			sh.tmp1[ro + n_new] = r2, sh.sidx[ro + n_new] = i; // parked; merged into place below
			++n_new;
		}
	}
	for (int t = n_new - 1, dst = n + n_new - 1, src = n - 1; t >= 0; --t) { // backward merge: tail t goes right behind the hit it was cut from
		for (; src > sh.sidx[ro + t]; --dst, --src) sh.r1[ro + dst] = sh.r1[ro + src], sh.pl[ro + dst] = sh.pl[ro + src];
		sh.r1[ro + dst] = sh.tmp1[ro + t], sh.pl[ro + dst].state = 0;
		--dst;
	}
	sh.n_reg[read] = n + n_new;
	if (n_new) {
#ifdef __CUDA_ARCH__
		atomicAdd(sh.ctr, (unsigned int)n_new);
#else
		sh.ctr[0] += (unsigned int)n_new;
#endif
	}
}

// tail of mm_align_skeleton (align.c:909-912) and of align_regs (map.c:264-268)
MMG_HDN inline void post_final(const PostShard &sh, int read)
{
	int n = sh.n_reg[read];
	if (n == 0) return;
	const int64_t ro = sh.roff[read];
	HitRec *r = sh.r1 + ro;
	RsFrame *st = sh.sstack + ro / 65 + 2 * (int64_t)read;
	n = hit_filter(sh.opt, sh.seq_len[read], n, r, sh.xw);
	n = hit_final_sort(n, r, sh.xw, sh.skey + ro, sh.sidx + ro, sh.tmp1 + ro, sh.sbig + ro, st);
	if (!(sh.opt.flag & HIT_F_ALL_CHAINS)) {
		hit_set_parent(sh.opt.mask_level, n, r, sh.opt.a * 2 + sh.opt.b, (sh.opt.flag & HIT_F_HARD_MLEVEL) != 0, sh.sidx + ro, sh.skey + ro, sh.xw);
		n = hit_select_sub(sh.opt.pri_ratio, sh.idx_k * 2, sh.opt.best_n, n, r, sh.sidx + ro);
		hit_set_sam_pri(n, r);
	}
	sh.n_reg[read] = n;
}

// mm_set_mapq per read, mm_pair, and the coordinate flip of mates that were mapped reverse-complemented (map.c:392-406, 486-497)
MMG_HDN inline void post_finish(const PostShard &sh, int f)
{
	const int ns = sh.n_seg[f], off = sh.seg_off[f], qlen_sum = post_qlen_sum(sh, f);
	if (!post_frag_mapped(sh, f, qlen_sum)) return;
	bool ok = true;
	for (int j = 0; j < ns; ++j)
		ok &= hit_set_mapq(sh.n_reg[off + j], sh.r1 + sh.roff[off + j], sh.xw, sh.opt.min_chain_score, sh.opt.a, sh.rep[f], (sh.opt.flag & HIT_F_SR) != 0, sh.lt, sh.skey + sh.roff[off + j]);
	if (ns == 2 && sh.opt.pe_ori >= 0 && (sh.opt.flag & HIT_F_CIGAR)) {
		int32_t ql[2] = {sh.seq_len[off], sh.seq_len[off + 1]}, nr[2] = {sh.n_reg[off], sh.n_reg[off + 1]};
		HitRec *rr[2] = {sh.r1 + sh.roff[off], sh.r1 + sh.roff[off + 1]};
		// scratch: the two reads' slots are adjacent (cap entries each), enough for n0 + n1 elements
		const int64_t ro = sh.roff[off];
		ok &= hit_pair(post_frag_gap(sh.opt, qlen_sum), sh.opt.pe_bonus, sh.opt.a * 2 + sh.opt.b, sh.opt.a, ql, nr, rr, sh.xw, sh.lt, sh.skey + ro, sh.sidx + ro,
		               reinterpret_cast<int32_t*>(sh.tmp1 + ro), sh.sbig + ro, sh.sstack + ro / 65 + 2 * (int64_t)off);
	}
	for (int j = 0; j < ns; ++j)
		if (sh.flip[off + j]) {
			HitRec *r = sh.r1 + sh.roff[off + j];
			const int ql = sh.seq_len[off + j];
			for (int k = 0; k < sh.n_reg[off + j]; ++k) { const int t = r[k].qs; r[k].qs = ql - r[k].qe, r[k].qe = ql - t, r[k].bits ^= HB_REV; }
		}
	if (!ok) {
#ifdef __CUDA_ARCH__
		atomicOr(sh.ctr + 1, POST_ERR_LOGTAB);
#else
		sh.ctr[1] |= POST_ERR_LOGTAB;
#endif
	}
}
