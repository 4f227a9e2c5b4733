// mmg_post.cu -- everything between chaining and MAPQ on the device, for the short-read presets.
//
// The reference does this bookkeeping per fragment inside mm_map_frag (map.c:376-400): chains -> mm_reg1_t
// (hit.c), primary/secondary selection, per-mate split (mm_seg_gen), the alignment skeleton with its DP calls
// (align.c), CIGAR post-processing, filters.  The host build of this repository runs the same steps from
// airlift_b200/host/{hits,aln}.c on the worker threads, which made the step host-bound.
// Here the SAME C SOURCES are compiled for the device (MM_DEVICE_BUILD, MM_FN = __device__) and run one fragment per
// thread, with the resumable DP-request mechanism of aln.c feeding K4 directly in device memory:
//
//     k_post_hits      chains -> hits -> per-mate regions -> first walk (queues the DP jobs)
//   { k_post_gather    jobs of all fragments -> one dense array
//     K4               mmg_ksw_device (results stay on the device)
//     k_post_align   } results -> job caches, re-walk (CIGAR stitching, mm_update_extra, filters, sorting)
//     k_post_sizes/pack  mm_reg1_t + mm_extra_t of every read -> one blob -> one D2H
//
// MAPQ and pairing (logf), the mate un-flip and the malloc'd API objects are made by the host from the blob.
// Memory for the per-fragment state comes from a bump pool in HBM (one atomicAdd per 2 KB chunk per thread).
#include <cub/cub.cuh>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
#include <assert.h>
#include <ctype.h>
#include <pthread.h>
#include "mmg_ctx.cuh"

int mmg_ksw_device(mmg_ctx_t *c, const mmg_idx_t *mi, const mmg_mapopt_t *opt, int n_jobs, const mmg_ksw_job_t *d_jobs, const void **d_res,
                   const uint32_t **d_cigars, double *kernel_ms, uint64_t *cells);

// The mapper's types and C sources live in a namespace of their own in this translation unit: the same declarations exist
// as host functions in the library, and the device twins must not be confused with them.
namespace mmdev {
#define MM_FN __device__
#define MM_DEVICE_BUILD 1
#include "../host/mm2b_priv.h"

struct KswResD { mmg_extz_t ez; uint64_t cigar_off; }; // == KswResDev of mmg_ksw.cu

struct Tls { char *cur, *end; };

struct DFrag {
	int n_segs, qlen_sum, frag_gap, active, n_regs0, n_u;
	uint32_t hash;
	int *qlens;
	mm128_t *a;
	uint64_t *u;
	mm_reg1_t *regs0;
	mm_seg_t *seg;
	mm_alnseg_t *aln;
};

struct Shard {
	const mm_idx_t *mi;
	mm_mapopt_t opt;
	int nf, n_seq, nt;                          // nt: threads of the per-fragment kernels (nf rounded up to whole warps)
	const int32_t *n_seg, *seg_off, *seq_len;   // per fragment / per read of the resident batch
	const uint64_t *q_off;
	const uint32_t *Q;
	const int32_t *nu, *nv;                     // chains per fragment (dense outputs of mmg_seed_chain_resident)
	const int64_t *uoff, *voff;
	uint64_t *u;
	mm128_t *a;
	const uint32_t *hash;
	mm_reg1_t **pre_regs;                       // per fragment: mm_gen_regs output made by k_post_regs_heavy (or null)
	const int32_t *perm;                        // thread -> fragment: fragments with similar amounts of work share a warp
	int32_t *n_reg;                             // per read
	mm_reg1_t **reg;
	DFrag *fr;
	Tls *tls;
	char *pool;
	unsigned long long pool_size;
	unsigned long long *pool_cur;               // [0] bump cursor, [1] active fragments, [2] error flag, [3] debug: longest thread (cycles << 20 | chains)
	int32_t *n_new;                             // per fragment: DP jobs queued by the last walk
	const int64_t *job_off;
	mmg_ksw_job_t *jobs;
	const KswResD *res;
	const uint32_t *cig;
};

// ---- device runtime the C sources run on --------------------------------------------------------------------
extern __shared__ __align__(16) unsigned char dyn_smem[];
__device__ __forceinline__ Shard *cur_shard() { return *reinterpret_cast<Shard**>(dyn_smem); }
__device__ __forceinline__ int cur_tid() { return blockIdx.x * blockDim.x + threadIdx.x; }

#define POOL_CHUNK 4096

__device__ void *dev_malloc(size_t n)
{ // 16-byte header (the size, for realloc) + payload, bumped inside the thread's current chunk
	Shard *sh = cur_shard();
	Tls *t = &sh->tls[cur_tid()];
	const size_t need = ((n + 15) & ~(size_t)15) + 16;
	if (t->cur == nullptr || t->cur + need > t->end) {
		const size_t chunk = need > POOL_CHUNK ? need : POOL_CHUNK;
		const unsigned long long o = atomicAdd(sh->pool_cur, (unsigned long long)chunk);
		if (o + chunk > sh->pool_size) { // out of pool: flag it; the host reports and stops (no CPU fallback hides it)
			atomicExch(sh->pool_cur + 2, 1ULL);
			asm volatile("trap;");
		}
		if (need > POOL_CHUNK) { // a large block gets its own chunk; the current one keeps filling
			char *p = sh->pool + o;
			*reinterpret_cast<size_t*>(p) = n;
			return p + 16;
		}
		t->cur = sh->pool + o, t->end = t->cur + chunk;
	}
	char *p = t->cur;
	t->cur += need;
	*reinterpret_cast<size_t*>(p) = n;
	return p + 16;
}
// memcpy / memset / memmove of the C sources: pool blocks are 16-byte aligned, CIGARs and records are word arrays, so
// almost every call can move 16 or 4 bytes per step instead of the byte loop the compiler emits for unknown alignment
__device__ void *dev_memcpy(void *dst, const void *src, size_t n)
{
	const size_t a = reinterpret_cast<size_t>(dst) | reinterpret_cast<size_t>(src);
	size_t i = 0;
	if ((a & 15) == 0) { uint4 *d = static_cast<uint4*>(dst); const uint4 *s = static_cast<const uint4*>(src); for (; i + 16 <= n; i += 16) d[i >> 4] = s[i >> 4]; }
	if ((a & 3) == 0) { uint32_t *d = static_cast<uint32_t*>(dst); const uint32_t *s = static_cast<const uint32_t*>(src); for (; i + 4 <= n; i += 4) d[i >> 2] = s[i >> 2]; }
	for (; i < n; ++i) static_cast<unsigned char*>(dst)[i] = static_cast<const unsigned char*>(src)[i];
	return dst;
}
__device__ void *dev_memset(void *dst, int v, size_t n)
{
	const size_t a = reinterpret_cast<size_t>(dst);
	const uint32_t w = 0x01010101u * (uint32_t)(unsigned char)v;
	size_t i = 0;
	if ((a & 15) == 0) { uint4 *d = static_cast<uint4*>(dst); const uint4 q = make_uint4(w, w, w, w); for (; i + 16 <= n; i += 16) d[i >> 4] = q; }
	if ((a & 3) == 0) { uint32_t *d = static_cast<uint32_t*>(dst); for (; i + 4 <= n; i += 4) d[i >> 2] = w; }
	for (; i < n; ++i) static_cast<unsigned char*>(dst)[i] = (unsigned char)v;
	return dst;
}
__device__ void *dev_memmove(void *dst, const void *src, size_t n)
{
	unsigned char *d = static_cast<unsigned char*>(dst); const unsigned char *s = static_cast<const unsigned char*>(src);
	if (d <= s || d >= s + n) return dev_memcpy(dst, src, n); // a forward copy never overwrites unread source bytes
	if (((reinterpret_cast<size_t>(d) | reinterpret_cast<size_t>(s) | n) & 7) == 0) {
		uint64_t *dd = static_cast<uint64_t*>(dst); const uint64_t *ss = static_cast<const uint64_t*>(src);
		for (size_t i = n >> 3; i > 0; --i) dd[i - 1] = ss[i - 1];
	} else for (size_t i = n; i > 0; --i) d[i - 1] = s[i - 1];
	return dst;
}
__device__ void *dev_calloc(size_t n, size_t sz) { void *p = dev_malloc(n * sz); dev_memset(p, 0, n * sz); return p; }
__device__ void *dev_realloc(void *p, size_t n)
{
	void *q = dev_malloc(n);
	if (p) { const size_t old = *reinterpret_cast<size_t*>(static_cast<char*>(p) - 16); dev_memcpy(q, p, old < n ? old : n); }
	return q;
}

// the arena interface of mm2b_priv.h, on the pool
__device__ void *mm_amalloc(size_t n) { return dev_malloc(n); }
__device__ void *mm_acalloc(size_t n, size_t sz) { return dev_calloc(n, sz); }
__device__ void *mm_arealloc(void *p, size_t old_bytes, size_t new_bytes) { void *q = dev_malloc(new_bytes); if (p && old_bytes) dev_memcpy(q, p, old_bytes < new_bytes ? old_bytes : new_bytes); return q; }
__device__ void mm_afree(void *) {}

__device__ int mm_idx_getseq_dev(const mm_idx_t *mi, uint32_t rid, uint32_t st, uint32_t en, uint8_t *seq)
{ // index.c:152-162 on the device copy of the index header (S, seq[] are device pointers)
	if (rid >= mi->n_seq || st >= mi->seq[rid].len) return -1;
	if (en > mi->seq[rid].len) en = mi->seq[rid].len;
	const uint64_t st1 = mi->seq[rid].offset + st, en1 = mi->seq[rid].offset + en;
	uint64_t i = st1;
	for (; i < en1 && (i & 7); ++i) seq[i - st1] = mm_seq4_get(mi->S, i);
	for (; i + 8 <= en1; i += 8) { // one packed word = 8 bases
		uint32_t w = mi->S[i >> 3];
		uint8_t *o = seq + (i - st1);
#pragma unroll
		for (int k = 0; k < 8; ++k) o[k] = (uint8_t)(w >> (4 * k) & 0xf);
	}
	for (; i < en1; ++i) seq[i - st1] = mm_seq4_get(mi->S, i);
	return (int)(en - st);
}

// ---- the C sources, compiled for the device -------------------------------------------------------------------
#define malloc(n) mmdev::dev_malloc(n)
#define calloc(n, s) mmdev::dev_calloc(n, s)
#define realloc(p, n) mmdev::dev_realloc(p, n)
#define free(p) ((void)(p))
#define memmove(d, s, n) mmdev::dev_memmove(d, s, n)
#define memcpy(d, s, n) mmdev::dev_memcpy(d, s, n)
#define memset(d, v, n) mmdev::dev_memset(d, v, n)
#define fprintf(...) ((void)0)
#define exit(c) asm volatile("trap;")
#define mm_verbose 0
#define mm_idx_getseq mm_idx_getseq_dev /* the header's name has C linkage and belongs to the host library */
#undef assert
#define assert(x) ((void)0)

RADIX_IMPL(radix_sort_128x, mm128_t, RS_KEY_X, RS_TABLES_ARENA)
RADIX_IMPL(radix_sort_64, uint64_t, RS_KEY_ID, RS_TABLES_ARENA)

#include "../host/llsw.c"
#include "../host/hits.c"
#include "../host/aln.c"

#undef malloc
#undef calloc
#undef realloc
#undef free
#undef memmove
#undef memcpy
#undef memset
#undef fprintf
#undef exit
#undef mm_verbose
#undef mm_idx_getseq

// ---- per-fragment stages (the device twins of mapper.c's stage_hits / stage_align / stage_scatter_results) -----

__device__ __forceinline__ int chain_gap_ref(const mm_mapopt_t *opt, int qlen_sum)
{ // map.c:344-349
	if (opt->max_gap_ref > 0) return opt->max_gap_ref;
	if (opt->max_frag_len > 0) { const int g = opt->max_frag_len - qlen_sum; return g < opt->max_gap ? opt->max_gap : g; }
	return opt->max_gap;
}

// one resumable pass over the regions of an active fragment (align_regs, map.c:260-270); returns the jobs it queued
__device__ int frag_walk(Shard *sh, int i)
{
	DFrag *fr = &sh->fr[i];
	const mm_mapopt_t *opt = &sh->opt;
	int all_done = 1, n_new = 0;
	if (!fr->active) return 0;
	for (int j = 0; j < fr->n_segs; ++j) {
		mm_alnseg_t *s = &fr->aln[j];
		const int off = sh->seg_off[i] + j;
		if (s->finished) continue;
		if (mm_aln_step(s, opt, sh->mi)) {
			if (!(opt->flag & MM_F_ALL_CHAINS)) {
				mm_set_parent(opt->mask_level, s->n_regs, s->regs, opt->a * 2 + opt->b, opt->flag & MM_F_HARD_MLEVEL);
				mm_select_sub(opt->pri_ratio, sh->mi->k * 2, opt->best_n, &s->n_regs, s->regs);
				mm_set_sam_pri(s->n_regs, s->regs);
			}
			sh->n_reg[off] = s->n_regs, sh->reg[off] = s->regs;
		} else all_done = 0;
		n_new += s->cache.n - s->cache.n_sent;
	}
	if (all_done) fr->active = 0;
	return n_new;
}

__global__ void __launch_bounds__(128, 5) k_post_hits(Shard *shp)
{ // chains -> hits for one fragment per thread (map.c:376-388, 390-400 up to the DP), then the first walk
	if (threadIdx.x == 0) *reinterpret_cast<Shard**>(dyn_smem) = shp;
	__syncthreads();
	Shard *sh = shp;
	if (cur_tid() >= sh->nt) return;
	const int i = sh->perm[cur_tid()];
	if (i < 0) return;
	const long long t_begin = clock64();
	const mm_mapopt_t *opt = &sh->opt;
	const mm_idx_t *mi = sh->mi;
	const int off = sh->seg_off[i], ns = sh->n_seg[i];
	DFrag *fr = &sh->fr[i];
	sh->tls[cur_tid()].cur = sh->tls[cur_tid()].end = nullptr;
	memset(fr, 0, sizeof(*fr));
	sh->n_new[i] = 0;
	fr->n_segs = ns;
	fr->qlens = (int*)mm_amalloc((size_t)(ns > 0 ? ns : 1) * sizeof(int));
	for (int j = 0; j < ns; ++j) {
		sh->n_reg[off + j] = 0, sh->reg[off + j] = nullptr;
		fr->qlens[j] = sh->seq_len[off + j], fr->qlen_sum += fr->qlens[j];
	}
	if (fr->qlen_sum == 0 || ns <= 0 || ns > MM_MAX_SEG) return;
	if (opt->max_qlen > 0 && fr->qlen_sum > opt->max_qlen) return;
	fr->hash = sh->hash[i];
	fr->frag_gap = chain_gap_ref(opt, fr->qlen_sum);
	fr->n_u = sh->nu[i];
	if (fr->n_u > 0) fr->u = sh->u + sh->uoff[i], fr->a = sh->a + sh->voff[i]; // used in place
	fr->regs0 = sh->pre_regs[i] ? sh->pre_regs[i] : mm_gen_regs(fr->hash, fr->qlen_sum, fr->n_u, fr->u, fr->a);
	fr->n_regs0 = fr->n_u;
	const bool dbg_big = fr->n_u > 1000;
	long long t_ph = clock64();
#define PHASE(k) do { if (dbg_big) { const long long t_ = clock64(); atomicMax(sh->pool_cur + 5 + (k), (unsigned long long)(t_ - t_ph)); t_ph = t_; } } while (0)
	PHASE(0);
	if (!(opt->flag & MM_F_ALL_CHAINS)) { // chain_post, map.c:249-258
		mm_set_parent(opt->mask_level, fr->n_regs0, fr->regs0, opt->a * 2 + opt->b, opt->flag & MM_F_HARD_MLEVEL);
		PHASE(1);
		if (ns <= 1) mm_select_sub(opt->pri_ratio, mi->k * 2, opt->best_n, &fr->n_regs0, fr->regs0);
		else mm_select_sub_multi(opt->pri_ratio, 0.2f, 0.7f, fr->frag_gap, mi->k * 2, opt->best_n, ns, fr->qlens, &fr->n_regs0, fr->regs0);
		if (!(opt->flag & (MM_F_SPLICE|MM_F_SR|MM_F_NO_LJOIN))) mm_join_long(opt, fr->qlen_sum, &fr->n_regs0, fr->regs0, fr->a);
	}
	PHASE(2);
	fr->aln = (mm_alnseg_t*)mm_acalloc(ns, sizeof(mm_alnseg_t));
	if (ns == 1) {
		sh->n_reg[off] = fr->n_regs0, sh->reg[off] = fr->regs0;
		mm_aln_begin(&fr->aln[0], off, fr->qlens[0], nullptr, fr->n_regs0, fr->regs0, fr->a);
		fr->aln[0].q4 = sh->Q, fr->aln[0].q4_off = sh->q_off[off];
	} else {
		fr->seg = mm_seg_gen(fr->hash, ns, fr->qlens, fr->n_regs0, fr->regs0, &sh->n_reg[off], &sh->reg[off], fr->a);
		for (int j = 0; j < ns; ++j) {
			mm_set_parent(opt->mask_level, sh->n_reg[off + j], sh->reg[off + j], opt->a * 2 + opt->b, opt->flag & MM_F_HARD_MLEVEL);
			mm_aln_begin(&fr->aln[j], off + j, fr->qlens[j], nullptr, sh->n_reg[off + j], sh->reg[off + j], fr->seg[j].a);
			fr->aln[j].q4 = sh->Q, fr->aln[j].q4_off = sh->q_off[off + j];
		}
	}
	PHASE(3);
	fr->active = (opt->flag & MM_F_CIGAR) ? 1 : 0;
	if (fr->active) {
		const int n_new = frag_walk(sh, i);
		sh->n_new[i] = n_new;
		if (fr->active) atomicAdd(sh->pool_cur + 1, 1ULL);
	}
	PHASE(4);
#undef PHASE
	{ // debug aid (MMG_POST_DEBUG): the longest-running thread of the launch and how many chains its fragment had
		const unsigned long long dt = (unsigned long long)(clock64() - t_begin);
		atomicMax(sh->pool_cur + 3, dt << 20 | (unsigned long long)(fr->n_u < 0xfffff ? fr->n_u : 0xfffff));
	}
}

__global__ void __launch_bounds__(128) k_post_gather(Shard *shp)
{ // copy the DP jobs a fragment queued in this round into the dense job array
	Shard *sh = shp;
	if (cur_tid() >= sh->nt) return;
	const int i = sh->perm[cur_tid()];
	if (i < 0) return;
	const DFrag *fr = &sh->fr[i];
	if (!fr->active) return;
	int64_t k = sh->job_off[i];
	for (int j = 0; j < fr->n_segs; ++j) {
		const mm_dpcache_t *c = &fr->aln[j].cache;
		for (int q = c->n_sent; q < c->n; ++q) sh->jobs[k++] = c->a[q].job;
	}
}

__global__ void __launch_bounds__(128, 5) k_post_align(Shard *shp)
{ // hand the results of this round back to the job caches (in the order they were gathered), then re-walk
	if (threadIdx.x == 0) *reinterpret_cast<Shard**>(dyn_smem) = shp;
	__syncthreads();
	Shard *sh = shp;
	if (cur_tid() >= sh->nt) return;
	const int i = sh->perm[cur_tid()];
	if (i < 0) return;
	DFrag *fr = &sh->fr[i];
	if (!fr->active) { sh->n_new[i] = 0; return; }
	int64_t k = sh->job_off[i];
	for (int j = 0; j < fr->n_segs; ++j) {
		mm_dpcache_t *c = &fr->aln[j].cache;
		for (; c->n_sent < c->n; ++c->n_sent, ++k) {
			mm_dpjob_t *dj = &c->a[c->n_sent];
			dj->ez = sh->res[k].ez;
			if (dj->ez.n_cigar > 0) {
				dj->cigar = (uint32_t*)mm_amalloc((size_t)dj->ez.n_cigar * 4);
				const uint32_t *src = sh->cig + sh->res[k].cigar_off;
				for (int q = 0; q < dj->ez.n_cigar; ++q) dj->cigar[q] = src[q];
			}
			dj->done = 1;
		}
	}
	const int n_new = frag_walk(sh, i);
	sh->n_new[i] = n_new;
	if (fr->active) atomicAdd(sh->pool_cur + 1, 1ULL);
}


// ---- mm_gen_regs (hit.c:52-88) for fragments with many chains, one CTA per fragment ---------------------------------
// A read pair from a high-copy repeat arrives with thousands of chains over ~10^5 anchors; walked by the single thread
// of k_post_hits it set the duration of the whole launch (34 ms).  The sort key (score, salted hash of the first
// anchor) is unique per chain unless two 32-bit hashes collide, so any sort gives klib's order; a collision is detected
// and leaves the fragment to the literal serial code.
#define POST_HEAVY_NU 128

__global__ void k_post_heavy_list(int nf, const int32_t *nu, int32_t *list, unsigned long long *counter)
{
	const int i = cur_tid();
	if (i < nf && nu[i] > POST_HEAVY_NU) list[atomicAdd(counter, 1ULL)] = i;
}

__global__ void __launch_bounds__(256) k_post_regs_heavy(Shard *shp, const int32_t *list)
{
	Shard *sh = shp;
	const int i = list[blockIdx.x], n = sh->nu[i], tid = threadIdx.x, nt = blockDim.x;
	const uint64_t *u = sh->u + sh->uoff[i];
	const mm128_t *a = sh->a + sh->voff[i];
	__shared__ char *s_base;
	__shared__ int s_scan[256], s_carry, s_tie;
	const size_t z_bytes = ((size_t)n * 16 + 15) & ~(size_t)15, r_bytes = (((size_t)n * sizeof(mm_reg1_t)) + 15) & ~(size_t)15;
	if (tid == 0) {
		const unsigned long long o = atomicAdd(sh->pool_cur, (unsigned long long)(z_bytes + r_bytes + 32));
		if (o + z_bytes + r_bytes + 32 > sh->pool_size) { atomicExch(sh->pool_cur + 2, 1ULL); asm volatile("trap;"); }
		s_base = sh->pool + o;
		s_carry = 0, s_tie = 0;
	}
	__syncthreads();
	mm128_t *z = reinterpret_cast<mm128_t*>(s_base + 16);
	char *rblk = s_base + 16 + z_bytes;
	mm_reg1_t *r = reinterpret_cast<mm_reg1_t*>(rblk + 16);
	if (tid == 0) *reinterpret_cast<size_t*>(s_base) = (size_t)n * 16, *reinterpret_cast<size_t*>(rblk) = (size_t)n * sizeof(mm_reg1_t); // block headers (realloc reads them)
	int qlen = 0;
	for (int j = 0; j < sh->n_seg[i]; ++j) qlen += sh->seq_len[sh->seg_off[i] + j];
	// first anchor of every chain: exclusive scan of the chain lengths, 256 at a time
	for (int i0 = 0; i0 < n; i0 += nt) {
		const int e = i0 + tid, cnt = e < n ? (int32_t)u[e] : 0;
		s_scan[tid] = cnt;
		__syncthreads();
		for (int d = 1; d < nt; d <<= 1) { const int v = tid >= d ? s_scan[tid - d] : 0; __syncthreads(); s_scan[tid] += v; __syncthreads(); }
		const int k = s_carry + s_scan[tid] - cnt;
		if (e < n) {
			const uint32_t h = (uint32_t)mix64((mix64(a[k].x) + mix64(a[k].y)) ^ sh->hash[i]);
			z[e].x = u[e] ^ h, z[e].y = (uint64_t)k << 32 | (uint32_t)(int32_t)u[e];
		}
		__syncthreads();
		if (tid == nt - 1) s_carry += s_scan[tid];
		__syncthreads();
	}
	// descending by x: bitonic network oriented one way, so the virtual entries beyond n never move
	int N = 2; while (N < n) N <<= 1;
	for (int k = 2; k <= N; k <<= 1) {
		for (int e = tid; e < n; e += nt) { const int l = e ^ (k - 1); if (l > e && l < n) { const mm128_t x = z[e], y = z[l]; if (y.x > x.x) z[e] = y, z[l] = x; } }
		__syncthreads();
		for (int j = k >> 2; j > 0; j >>= 1) {
			for (int e = tid; e < n; e += nt) { const int l = e ^ j; if (l > e && l < n) { const mm128_t x = z[e], y = z[l]; if (y.x > x.x) z[e] = y, z[l] = x; } }
			__syncthreads();
		}
	}
	for (int e = tid + 1; e < n; e += nt) if (z[e].x == z[e - 1].x) s_tie = 1;
	__syncthreads();
	if (s_tie) return; // equal keys: klib's unstable order matters, k_post_hits runs the literal mm_gen_regs
	for (int e = tid; e < n; e += nt) {
		mm_reg1_t *ri = &r[e];
		memset(ri, 0, sizeof(*ri));
		ri->id = e, ri->parent = MM_PARENT_UNSET;
		ri->score = ri->score0 = (int32_t)(z[e].x >> 32);
		ri->hash = (uint32_t)z[e].x;
		ri->cnt = (int32_t)z[e].y, ri->as = (int32_t)(z[e].y >> 32);
		ri->div = -1.0f;
		reg_set_coor(ri, qlen, a);
	}
	if (tid == 0) sh->pre_regs[i] = r;
}

// ---- results: every read's mm_reg1_t array followed by the mm_extra_t of each hit, 8-byte aligned pieces
__device__ __forceinline__ size_t extra_bytes(const mm_extra_t *p) { return (sizeof(mm_extra_t) + (size_t)p->n_cigar * 4 + 7) & ~(size_t)7; }

__global__ void k_post_sizes(Shard *shp, int64_t *sizes)
{
	Shard *sh = shp;
	const int r = cur_tid();
	if (r > sh->n_seq) return;
	int64_t b = 0;
	if (r < sh->n_seq) {
		const int n = sh->n_reg[r];
		b = (int64_t)n * (int64_t)sizeof(mm_reg1_t);
		for (int i = 0; i < n; ++i) if (sh->reg[r][i].p) b += (int64_t)extra_bytes(sh->reg[r][i].p);
	}
	sizes[r] = b;
}

__global__ void k_post_pack(Shard *shp, const int64_t *offs, unsigned char *blob)
{
	Shard *sh = shp;
	const int r = cur_tid();
	if (r >= sh->n_seq) return;
	const int n = sh->n_reg[r];
	unsigned char *o = blob + offs[r];
	const mm_reg1_t *regs = sh->reg[r];
	for (int i = 0; i < n; ++i) reinterpret_cast<mm_reg1_t*>(o)[i] = regs[i];
	o += (size_t)n * sizeof(mm_reg1_t);
	for (int i = 0; i < n; ++i)
		if (regs[i].p) {
			const size_t b = sizeof(mm_extra_t) + (size_t)regs[i].p->n_cigar * 4;
			const uint32_t *src = reinterpret_cast<const uint32_t*>(regs[i].p);
			uint32_t *dst = reinterpret_cast<uint32_t*>(o);
			for (size_t w = 0; w < b / 4; ++w) dst[w] = src[w];
			o += extra_bytes(regs[i].p);
		}
}

__global__ void k_post_keys(int nf, const int32_t *nu, const int32_t *nv, uint32_t *key, int32_t *idx)
{ // work estimate of a fragment: its chains, then its chained anchors (descending, so the heavy warps start first)
	const int i = cur_tid();
	if (i >= nf) return;
	const uint32_t u = nu[i] < 1023 ? (uint32_t)nu[i] : 1023u, v = nv[i] < 0xfffff ? (uint32_t)nv[i] : 0xfffffu;
	key[i] = ~(u << 20 | v), idx[i] = i;
}

__global__ void k_post_spread(int nf, int n_warps, const int32_t *sorted, int32_t *perm)
{ // rank r (heaviest first) -> warp r % n_warps, lane r / n_warps: every warp gets one fragment of each weight class, the
  // heaviest fragments sit alone at the front of the launch instead of serialising inside one warp
	const int t = cur_tid();
	if (t >= n_warps * 32) return;
	const int64_t r = (int64_t)(t & 31) * n_warps + (t >> 5);
	perm[t] = r < nf ? sorted[r] : -1;
}

__global__ void k_post_mi(mm_idx_t *mi, mm_idx_seq_t *seq, int n_seq, const uint64_t *seq_off, const uint32_t *seq_len, uint32_t *S, int k, int w, int flag)
{ // device copy of the index header: what hits.c / aln.c read through mm_idx_t
	const int i = cur_tid();
	if (i < n_seq) seq[i].name = nullptr, seq[i].offset = seq_off[i], seq[i].len = seq_len[i];
	if (i == 0) {
		memset(mi, 0, sizeof(*mi));
		mi->k = k, mi->w = w, mi->flag = flag, mi->n_seq = (uint32_t)n_seq, mi->seq = seq, mi->S = S;
	}
}

} // namespace mmdev

using namespace mmdev;

template <class T> static int post_scan(mmg_ctx_t *c, const T *d_in, int64_t *d_out, int n)
{
	size_t tmp = 0;
	cub::DeviceScan::ExclusiveSum(nullptr, tmp, d_in, d_out, n, c->stream);
	MMG_TRY(c->d_cub.ensure(tmp));
	MMG_CUDA(cub::DeviceScan::ExclusiveSum(c->d_cub.p, tmp, d_in, d_out, n, c->stream));
	++c->launches;
	return MMG_OK;
}

// Runs the stages above for the batch resident on `c`, right after mmg_seed_chain_resident(download = 0).
extern "C" int mmg_post_chain(mmg_ctx_t *c, const mmg_idx_t *mi, const mmg_mapopt_t *dopt, const void *opt_full, size_t opt_bytes, int idx_flag,
                              const uint32_t *frag_hash, mmg_post_out_t *out)
{
	memset(out, 0, sizeof(*out));
	if (opt_bytes != sizeof(mm_mapopt_t)) { mmg_set_error("mmg_post_chain: mm_mapopt_t size mismatch"); return MMG_EINVAL; }
	MMG_CUDA(cudaSetDevice(c->dev));
	const ResidentBatch &rb = c->rb;
	const int nf = rb.n_frag, n_seq = rb.n_seq;
	if (nf == 0) return MMG_OK;
	{ // once per process: room for the per-thread call stacks of the compiled C code
		static bool stack_set[16] = {false};
		if (c->dev < 16 && !stack_set[c->dev]) { MMG_CUDA(cudaDeviceSetLimit(cudaLimitStackSize, 8192)); stack_set[c->dev] = true; }
	}
	if (c->p_mi_for != mi) { // device copy of the index header
		MMG_TRY(c->p_mi.ensure(sizeof(mm_idx_t) + 64));
		MMG_TRY(c->p_seq.ensure((size_t)(mi->n_seq + 1) * sizeof(mm_idx_seq_t)));
		MMG_LAUNCH(c, k_post_mi, mmg_blocks(mi->n_seq > 0 ? mi->n_seq : 1, 128), 128, 0, c->p_mi.as<mm_idx_t>(), c->p_seq.as<mm_idx_seq_t>(), mi->n_seq, mi->d_seq_off,
		           mi->d_seq_len, mi->d_S, mi->k, mi->w, idx_flag);
		c->p_mi_for = mi;
	}
	// measured use on 2x150 bp pairs: ~10 KB per fragment (MMG_POST_DEBUG prints it); the per-base term covers long queries
	const size_t pool_bytes = (size_t)nf * 24576 + (size_t)rb.n_bases * 32 + ((size_t)512 << 20);
	MMG_TRY(c->p_shard.ensure(sizeof(Shard)));
	MMG_TRY(c->p_hash.ensure((size_t)(nf + 1) * 4));
	MMG_TRY(c->p_nreg.ensure((size_t)(n_seq + 1) * 4));
	MMG_TRY(c->p_reg.ensure((size_t)(n_seq + 1) * 8));
	MMG_TRY(c->p_fr.ensure((size_t)(nf + 1) * sizeof(DFrag)));
	MMG_TRY(c->p_tls.ensure((size_t)(nf + 64) * sizeof(Tls)));
	MMG_TRY(c->p_pool.ensure(pool_bytes));
	MMG_TRY(c->p_ctr.ensure(128));
	MMG_TRY(c->p_nnew.ensure((size_t)(nf + 2) * 4));
	MMG_TRY(c->p_joboff.ensure((size_t)(nf + 2) * 8));
	MMG_TRY(c->h_p_hash.ensure((size_t)(nf + 1) * 4));
	memcpy(c->h_p_hash.p, frag_hash, (size_t)nf * 4);
	MMG_H2D(c, c->p_hash.p, c->h_p_hash.p, (size_t)nf * 4);
	MMG_CUDA(cudaMemsetAsync(c->p_ctr.p, 0, 128, c->stream));
	MMG_CUDA(cudaMemsetAsync(c->p_nnew.p, 0, (size_t)(nf + 2) * 4, c->stream));
	Shard hs;
	memset(&hs, 0, sizeof(hs));
	hs.mi = c->p_mi.as<mm_idx_t>();
	memcpy(&hs.opt, opt_full, sizeof(mm_mapopt_t));
	hs.opt.split_prefix = nullptr;
	hs.nf = nf, hs.n_seq = n_seq;
	hs.n_seg = c->d_misc.as<int32_t>(), hs.seg_off = nullptr, hs.seq_len = c->d_seq_len.as<int32_t>();
	hs.q_off = c->d_q_off.as<uint64_t>(), hs.Q = c->d_Q.as<uint32_t>();
	hs.nu = c->d_frag_nu.as<int32_t>(), hs.nv = c->d_frag_nv.as<int32_t>();
	hs.uoff = c->d_uoff.as<int64_t>(), hs.voff = c->d_voff.as<int64_t>();
	hs.u = c->d_out_u.as<uint64_t>(), hs.a = c->d_out_a.as<mm128_t>();
	hs.hash = c->p_hash.as<uint32_t>();
	hs.n_reg = c->p_nreg.as<int32_t>(), hs.reg = c->p_reg.as<mm_reg1_t*>();
	hs.fr = c->p_fr.as<DFrag>(), hs.tls = c->p_tls.as<Tls>();
	hs.pool = c->p_pool.as<char>(), hs.pool_size = pool_bytes, hs.pool_cur = c->p_ctr.as<unsigned long long>();
	hs.n_new = c->p_nnew.as<int32_t>(), hs.job_off = c->p_joboff.as<int64_t>();
	// first read of each fragment: the resident batch keeps it on the host; upload once per batch
	MMG_TRY(c->d_frag_list.ensure((size_t)(nf + 1) * 8));
	{
		// d_frag_list's upper half ([nf, 2nf)) holds the second-pass source slots until the gather; seg_off goes to its own buffer
		MMG_TRY(c->p_sizes.ensure((size_t)(n_seq + nf + 4) * 8));
		int32_t *d_segoff = reinterpret_cast<int32_t*>(c->p_sizes.as<int64_t>() + n_seq + 2);
		MMG_H2D(c, d_segoff, rb.seg_off.data(), (size_t)nf * 4);
		hs.seg_off = d_segoff;
	}
	{ // thread -> fragment order
		const int n_warps = (nf + 31) / 32;
		hs.nt = n_warps * 32;
		MMG_TRY(c->p_perm.ensure((size_t)(nf + 33) * 20));
		uint32_t *key = c->p_perm.as<uint32_t>(), *key2 = key + nf + 1;
		int32_t *idx = reinterpret_cast<int32_t*>(key2 + nf + 1), *sorted = idx + nf + 1, *perm = sorted + nf + 1;
		MMG_LAUNCH(c, k_post_keys, mmg_blocks(nf, 256), 256, 0, nf, hs.nu, hs.nv, key, idx);
		size_t tmp = 0;
		cub::DeviceRadixSort::SortPairs(nullptr, tmp, key, key2, idx, sorted, nf, 0, 32, c->stream);
		MMG_TRY(c->d_cub.ensure(tmp));
		MMG_CUDA(cub::DeviceRadixSort::SortPairs(c->d_cub.p, tmp, key, key2, idx, sorted, nf, 0, 32, c->stream));
		++c->launches;
		MMG_LAUNCH(c, k_post_spread, mmg_blocks(hs.nt, 256), 256, 0, nf, n_warps, sorted, perm);
		hs.perm = perm;
	}
	if (getenv("MMG_TRACE")) { // chains per fragment of this shard (debug aid; synchronises)
		std::vector<int32_t> hu((size_t)nf), hv((size_t)nf);
		cudaStreamSynchronize(c->stream);
		cudaMemcpy(hu.data(), hs.nu, (size_t)nf * 4, cudaMemcpyDeviceToHost);
		cudaMemcpy(hv.data(), hs.nv, (size_t)nf * 4, cudaMemcpyDeviceToHost);
		const int lim[8] = {0, 1, 2, 4, 16, 128, 1024, 1 << 30};
		long long cnt[8] = {0}, ch[8] = {0}, an[8] = {0};
		for (int i = 0; i < nf; ++i) { int b = 0; while (hu[i] > lim[b]) ++b; ++cnt[b], ch[b] += hu[i], an[b] += hv[i]; }
		fprintf(stderr, "[mmg::post] %d fragments; chains per fragment (fragments/chains/chained anchors):", nf);
		for (int b = 0; b < 8; ++b) fprintf(stderr, " <=%d: %lld/%lld/%lld", lim[b], cnt[b], ch[b], an[b]);
		fprintf(stderr, "\n");
	}
	MMG_TRY(c->p_pre.ensure((size_t)(nf + 1) * 12 + 64));
	hs.pre_regs = c->p_pre.as<mm_reg1_t*>();
	int32_t *d_heavy = reinterpret_cast<int32_t*>(hs.pre_regs + nf + 1);
	MMG_CUDA(cudaMemsetAsync(c->p_pre.p, 0, (size_t)(nf + 1) * 8, c->stream));
	MMG_H2D(c, c->p_shard.p, &hs, sizeof(hs));
	unsigned long long n_heavy = 0;
	MMG_LAUNCH(c, k_post_heavy_list, mmg_blocks(nf, 256), 256, 0, nf, hs.nu, d_heavy, c->p_ctr.as<unsigned long long>() + 4);
	MMG_D2H(c, &n_heavy, c->p_ctr.as<unsigned long long>() + 4, 8);
	MMG_CUDA(cudaStreamSynchronize(c->stream)); // hs lives on this stack frame
	MMG_CUDA(cudaEventRecord(c->ev[0], c->stream));
	if (n_heavy) MMG_LAUNCH(c, k_post_regs_heavy, (unsigned)n_heavy, 256, 0, c->p_shard.as<Shard>(), d_heavy);
	MMG_LAUNCH(c, k_post_hits, mmg_blocks(hs.nt, 128), 128, 16, c->p_shard.as<Shard>());
	double ksw_ms = 0;
	for (int round = 0;; ++round) {
		unsigned long long ctr[3] = {0, 0, 0};
		int64_t n_jobs = 0;
		MMG_TRY(post_scan(c, hs.n_new, c->p_joboff.as<int64_t>(), nf + 1));
		MMG_D2H(c, ctr, c->p_ctr.p, 24);
		MMG_D2H(c, &n_jobs, c->p_joboff.as<int64_t>() + nf, 8);
		cudaError_t e = cudaStreamSynchronize(c->stream);
		if (e != cudaSuccess) { mmg_set_error("post-chaining stages: %s%s", cudaGetErrorString(e), " (device pool exhausted or a kernel fault)"); return MMG_ECUDA; }
		if (ctr[1] == 0) break; // no active fragment left
		if (n_jobs == 0) { mmg_set_error("post-chaining stages: alignment made no progress"); return MMG_ECUDA; }
		if (n_jobs > 0x7fffffff) { mmg_set_error("post-chaining stages: too many DP jobs in one round"); return MMG_ELIMIT; }
		MMG_TRY(c->p_jobs.ensure((size_t)(n_jobs + 1) * sizeof(mmg_ksw_job_t)));
		hs.jobs = c->p_jobs.as<mmg_ksw_job_t>();
		MMG_H2D(c, reinterpret_cast<char*>(c->p_shard.p) + offsetof(Shard, jobs), &hs.jobs, sizeof(hs.jobs));
		MMG_LAUNCH(c, k_post_gather, mmg_blocks(hs.nt, 128), 128, 0, c->p_shard.as<Shard>());
		const void *d_res = nullptr; const uint32_t *d_cig = nullptr; double kms = 0; uint64_t cells = 0;
		MMG_TRY(mmg_ksw_device(c, mi, dopt, (int)n_jobs, hs.jobs, &d_res, &d_cig, &kms, &cells));
		ksw_ms += kms;
		out->n_dp_jobs += (uint64_t)n_jobs, out->n_dp_cells += cells, out->n_dp_rounds += 1;
		out->n_dp_jobs_fast += c->k_last_jobs_fast, out->n_dp_cells_fast += c->k_last_cells_fast;
		hs.res = static_cast<const KswResD*>(d_res), hs.cig = d_cig;
		MMG_H2D(c, reinterpret_cast<char*>(c->p_shard.p) + offsetof(Shard, res), &hs.res, sizeof(hs.res));
		MMG_H2D(c, reinterpret_cast<char*>(c->p_shard.p) + offsetof(Shard, cig), &hs.cig, sizeof(hs.cig));
		MMG_CUDA(cudaMemsetAsync(c->p_ctr.as<unsigned long long>() + 1, 0, 8, c->stream));
		MMG_LAUNCH(c, k_post_align, mmg_blocks(hs.nt, 128), 128, 16, c->p_shard.as<Shard>());
	}
	// pack and download
	int64_t *sizes = c->p_sizes.as<int64_t>();
	MMG_TRY(c->p_offs.ensure((size_t)(n_seq + 2) * 8));
	MMG_LAUNCH(c, k_post_sizes, mmg_blocks(n_seq + 1, 128), 128, 0, c->p_shard.as<Shard>(), sizes);
	MMG_TRY(post_scan(c, sizes, c->p_offs.as<int64_t>(), n_seq + 1));
	int64_t blob_bytes = 0;
	MMG_D2H(c, &blob_bytes, c->p_offs.as<int64_t>() + n_seq, 8);
	MMG_CUDA(cudaStreamSynchronize(c->stream));
	MMG_TRY(c->p_blob.ensure((size_t)blob_bytes + 64));
	MMG_LAUNCH(c, k_post_pack, mmg_blocks(n_seq, 128), 128, 0, c->p_shard.as<Shard>(), c->p_offs.as<int64_t>(), c->p_blob.as<unsigned char>());
	MMG_CUDA(cudaEventRecord(c->ev[1], c->stream));
	MMG_TRY(c->h_p_nreg.ensure((size_t)(n_seq + 1) * 4));
	MMG_TRY(c->h_p_offs.ensure((size_t)(n_seq + 2) * 8));
	MMG_TRY(c->h_p_blob.ensure((size_t)blob_bytes + 64));
	MMG_TRY(c->h_p_rep.ensure((size_t)(nf + 1) * 4));
	MMG_D2H(c, c->h_p_nreg.p, c->p_nreg.p, (size_t)n_seq * 4);
	MMG_D2H(c, c->h_p_offs.p, c->p_offs.p, (size_t)(n_seq + 1) * 8);
	if (blob_bytes) MMG_D2H(c, c->h_p_blob.p, c->p_blob.p, (size_t)blob_bytes);
	MMG_D2H(c, c->h_p_rep.p, c->d_frag_rep.p, (size_t)nf * 4);
	{
		cudaError_t e = cudaStreamSynchronize(c->stream);
		if (e != cudaSuccess) { mmg_set_error("post-chaining stages (pack): %s", cudaGetErrorString(e)); return MMG_ECUDA; }
	}
	float ms = 0;
	cudaEventElapsedTime(&ms, c->ev[0], c->ev[1]);
	if (getenv("MMG_POST_DEBUG")) {
		unsigned long long ctr4[10] = {0};
		cudaMemcpy(ctr4, c->p_ctr.p, 80, cudaMemcpyDeviceToHost);
		fprintf(stderr, "[mmg_post] phases of fragments with > 1000 chains, ms (max): gen_regs %.2f set_parent %.2f select_sub %.2f seg_gen+begin %.2f first walk %.2f\n",
		        ctr4[5] / 1.965e6, ctr4[6] / 1.965e6, ctr4[7] / 1.965e6, ctr4[8] / 1.965e6, ctr4[9] / 1.965e6);
		fprintf(stderr, "[mmg_post] %d fragments, pool used %.1f MB of %.1f MB (%.0f B per fragment), blob %.1f MB, %.1f ms on the device (K4 %.1f ms); "
		        "longest k_post_hits thread %.2f ms (fragment with %llu chains)\n", nf, ctr4[0] / 1e6, pool_bytes / 1e6, (double)ctr4[0] / nf, blob_bytes / 1e6, ms, ksw_ms,
		        (double)(ctr4[3] >> 20) / 1.965e6, ctr4[3] & 0xfffff);
	}
	out->n_reg = c->h_p_nreg.as<int32_t>(), out->blob_off = c->h_p_offs.as<int64_t>(), out->blob = c->h_p_blob.as<unsigned char>();
	out->rep_len = c->h_p_rep.as<int32_t>();
	out->t_device_ms = ms, out->t_ksw_ms = ksw_ms;
	return MMG_OK;
}
