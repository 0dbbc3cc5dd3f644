// poa_kernels.cuh -- sm_100a kernel of the POA read-vs-graph banded DP sweep.
//
// Replaces, for a BATCH of independent sweep jobs (one job = one call of the reference's align_rd_bspoacore):
//   align_rd_bspoacore                bspoa.h:2515-2618   stack-driven topological sweep + end candidates
//   dpalign_row_update_bspoa          bspoa.h:2232-2261   row_movx (bsalign.h:2244-2392) + piecex_row_cal (bsalign.h:2727-3185)
//   dpalign_row_merge_bspoa           bspoa.h:2263-2272   piecex_row_merge (bsalign.h:2474-2616)
//   row_init of the head node         bspoa.h:2224-2226   (bsalign.h:2094-2140)
//   the four query profiles           bspoa.h:2199-2215   replaced by a 1-byte code per query position (base | hpc flag << 2)
//
// Decomposition: one GROUP of 16 threads per job, thread j = SSE lane j of the reference (running block
// [j*W, (j+1)*W) of the band), scalar int arithmetic with the saturation bounds of the SSE byte ops, so every
// stored byte equals the reference's.  A job is inherently sequential over graph nodes (each row needs its
// predecessors' rows); parallelism comes from the 16 lanes and from many jobs in flight.
//
// Row blocks live in HBM in EXACTLY the reference's layout (dpalign_row_prepare_data, bspoa.h:1787-1793):
// node n of a job owns mmblk = roundup16(bw*(pw+1) + 68) bytes = [u bw][e bw][q bw][ubegs 17 x int32], bytes in
// striped order (band position p -> (p % W) * 16 + p / W), so the host can copy the arena straight into g->memp.
// Two row slots per group in shared memory (same striped layout) cache the predecessor row of the node being
// expanded and the row just computed: on the backbone of the graph the next predecessor is the row just written,
// so the HBM block is written once and not read back.
#pragma once
#include "common.cuh"

namespace bsb200 {

constexpr int kPoaGroup = 16;          // threads per job = SSE lanes
constexpr int kPoaThreads = 32;        // one warp per CTA = two jobs in flight per CTA
constexpr int kPoaMposInit = 0x7FFFFFFF - 1;   // MAX_B4 - 1, bspoa.h:2525

struct PoaArgs {
	uint32_t njobs;
	const uint32_t *order;       // job indices, heaviest first
	unsigned int *counter;
	const int32_t *par;          // njobs x 10: bandwidth, alnmode, M, X, O, E, Q, P, T, refbonus
	const uint8_t *qcode;        // per query position: base | (next base differs) << 2   (written by poa_prep_kernel)
	const uint64_t *qoff;
	const uint32_t *slen;
	const uint64_t *node_off;    // njobs + 1
	const int2 *node;            // per node: x = rpos, y = nct | base << 16 | bonus << 24
	const int32_t *eoff;         // per job nnode + 1 entries at node_off[job] + job
	const uint64_t *edge_off;    // njobs + 1
	const int32_t *edst;         // local node ids
	const uint32_t *head, *tail;
	int32_t *mpos;               // per node scratch
	uint32_t *vst;               // per node scratch (zeroed by poa_prep_kernel)
	uint32_t *stack;             // per node scratch
	uint8_t *rows;               // row block arena
	const uint64_t *row_off;     // per job byte offset into rows
	int32_t *best;               // njobs x 3: maxscr, maxidx (local node id), maxoff
	int32_t *status;             // njobs
	unsigned long long *ops;     // njobs x 2: row updates, row merges
	uint32_t slot_bytes;         // shared memory bytes per row slot (3 * max bw + 80)
};

// one block per job: query codes + traversal scratch
__global__ void poa_prep_kernel(uint32_t njobs, const uint32_t *order, const uint8_t *queries, const uint64_t *qoff, const uint32_t *slen, uint8_t *qcode,
		const uint64_t *node_off, int32_t *mpos, uint32_t *vst){
	for(uint32_t k=blockIdx.x;k<njobs;k+=gridDim.x){
		const uint32_t job = order[k];
		const uint64_t qo = qoff[job]; const uint32_t n = slen[job];
		for(uint32_t x=threadIdx.x;x<n;x+=blockDim.x){
			uint32_t b = queries[qo + x];
			uint32_t f = (x + 1 < n && queries[qo + x + 1] != b) ? 4u : 0u;   // set_query_prof_hpc, bsalign.h:2204-2206
			qcode[qo + x] = (uint8_t)((b & 3u) | f);
		}
		const uint64_t n0 = node_off[job], n1 = node_off[job + 1];
		for(uint64_t i=n0+threadIdx.x;i<n1;i+=blockDim.x){ mpos[i] = kPoaMposInit; vst[i] = 0; }
	}
}

__device__ __forceinline__ int sat16i(int v){ return max(-32768, min(32767, v)); }

// view of one row slot (shared memory) or row block: u/e/q striped, then 17 anchors
struct PoaSlot {
	int8_t *u, *e, *q; int32_t *ub;
	__device__ __forceinline__ void bind(uint8_t *p, uint32_t bw){ u = (int8_t*)p; e = u + bw; q = e + bw; ub = (int32_t*)(q + bw); }
};

__global__ void __launch_bounds__(kPoaThreads) poa_sweep_kernel(const PoaArgs a){
	extern __shared__ __align__(16) uint8_t poa_smem[];
	const int lane = threadIdx.x & 31;
	const int j = lane & 15;                               // SSE lane owned by this thread
	const unsigned gmask = 0xffffu << (lane & 16);
	const int gl0 = lane & 16;                             // first warp lane of this group
	// per group: two row slots, then 17 shifted anchors (+3 pad), 16 fend + 16 gend bytes, 32 ints for row_max
	uint8_t *gs = poa_smem + (size_t)(threadIdx.x >> 4) * (2 * a.slot_bytes + 80 + 32 + 128);
	uint8_t *slotA = gs, *slotB = gs + a.slot_bytes;
	int32_t *sUBS = (int32_t*)(gs + 2 * a.slot_bytes);
	int8_t *sF = (int8_t*)(sUBS + 20);
	int32_t *sRM = (int32_t*)(sF + 32);

	while(true){
		uint32_t idx = 0;
		if(j == 0) idx = atomicAdd(a.counter, 1u);
		idx = __shfl_sync(gmask, idx, gl0);
		if(idx >= a.njobs) break;
		const uint32_t job = a.order[idx];
		const int32_t *par = a.par + (size_t)job * 10;
		const uint32_t bw = (uint32_t)par[0], W = bw / kLanes;
		const int mode = par[1] & 3, Mm = par[2], Xx = par[3], O = par[4], E = par[5], Q = par[6], P = par[7], T = par[8], refbonus = par[9];
		const int go1 = (int8_t)O, ge1 = (int8_t)E, go2 = (int8_t)Q, ge2 = (int8_t)P;
		const int pw = epi8_piecewise(go1, ge1, go2, ge2, (int)bw);
		const int GOE = (int8_t)(go1 + ge1), GQP = (int8_t)(go2 + ge2), GOQ = clamp8(GOE - GQP);
		const int smax_nt = (int8_t)(Mm + refbonus + 1), smin_nt = (int8_t)Xx;        // bspoa.h:2226, 2240
		const uint32_t slen = a.slen[job];
		const uint8_t *qc = a.qcode + a.qoff[job];
		const uint64_t n0 = a.node_off[job];
		const int2 *node = a.node + n0;
		const int32_t *eoff = a.eoff + n0 + job;
		const int32_t *edst = a.edst + a.edge_off[job];
		int32_t *mpos = a.mpos + n0; uint32_t *vst = a.vst + n0; uint32_t *stack = a.stack + n0;
		const uint32_t head = a.head[job], tail = a.tail[job];
		const uint32_t mmblk = (bw * (pw + 1) + 68 + 15) / 16 * 16;               // bspoa.h:2217
		uint8_t *rows = a.rows + a.row_off[job];
		int maxscr = kScoreMin, maxidx = -1, maxoff = -1, stflag = 0;
		unsigned long long nupd = 0, nmrg = 0;
		// overhang constants of row_movx (bsalign.h:2357-2369)
		uint32_t ovd; int ovc;
		if(pw == 2){ ovd = (uint32_t)((go1 - go2) / (ge2 - ge1)); ovc = min(smin_nt, go2 + ge2) - 1 - smax_nt + (go2 + ge2); }
		else { ovd = bw + 1; ovc = min(smin_nt, go1 + ge1) - 1 - smax_nt + (go1 + ge1); }

		PoaSlot Pd, Cu;            // predecessor slot (node being expanded) and current slot (row just computed)
		Pd.bind(slotA, bw); Cu.bind(slotB, bw);
		uint32_t cur_node = 0xffffffffu;

		// slot -> HBM block of node n (reference layout), 16-byte pieces spread over the group
		auto store_block = [&](const PoaSlot &S, uint32_t n){
			uint8_t *dst = rows + (size_t)n * mmblk;
			const uint32_t npc = bw * (pw + 1) / 16;
			for(uint32_t c=j;c<npc;c+=kPoaGroup) *(uint4*)(dst + 16 * c) = *(const uint4*)((const uint8_t*)S.u + 16 * c);
			int32_t *dub = (int32_t*)(dst + bw * (pw + 1));
			dub[j] = S.ub[j];
			if(j == 0) dub[16] = S.ub[16];
		};
		auto load_block = [&](PoaSlot &S, uint32_t n){
			const uint8_t *src = rows + (size_t)n * mmblk;
			const uint32_t npc = bw * (pw + 1) / 16;
			for(uint32_t c=j;c<npc;c+=kPoaGroup) *(uint4*)((uint8_t*)S.u + 16 * c) = *(const uint4*)(src + 16 * c);
			const int32_t *sub = (const int32_t*)(src + bw * (pw + 1));
			S.ub[j] = sub[j];
			if(j == 0) S.ub[16] = sub[16];
		};
		// absolute H at band position pos of a slot (bsalign.h:3187-3197); every thread computes it
		auto getscore = [&](const PoaSlot &S, int64_t pos) -> int {
			if(pos < 0 || pos >= (int64_t)bw){ stflag |= 1; return kScoreMin; }
			uint32_t jj = (uint32_t)pos / W, ii = (uint32_t)pos - jj * W;
			int s = S.ub[jj];
			for(uint32_t k=0;k<=ii;k++) s += S.u[k * 16 + jj];
			return s;
		};

		// ---- head row (bsalign.h:2094-2140) ----------------------------------------------------------------
		{
			const bool two = (pw == 2), glob = (mode == 0 || mode == 2);
			const int ext = two ? ge2 : ge1;
			const int u0 = (int8_t)(go1 + ge1 + smin_nt - smax_nt);
			const uint32_t xp = two ? (uint32_t)((go2 - go1) / (ge1 - ge2)) : 0;
			for(uint32_t i=0;i<W;i++){
				uint32_t p = (uint32_t)j * W + i;
				int v = 0;
				if(glob) v = (p == 0) ? u0 : ((two && p < xp) ? ge1 : ext);
				Cu.u[i * 16 + j] = (int8_t)v;
				Cu.e[i * 16 + j] = (int8_t)(pw >= 1 ? kEpi8Min : 0);
				Cu.q[i * 16 + j] = (int8_t)(pw == 2 ? kEpi8Min : 0);
			}
			for(int k=j;k<=kLanes;k+=kPoaGroup){
				int s = 0;
				if(glob){
					int64_t n = (int64_t)k * W;
					s = smax_nt - smin_nt;
					if(n > 0){
						s += u0;
						int64_t n1 = 0;
						if(two){ n1 = (int64_t)xp - 1; if(n1 > n - 1) n1 = n - 1; if(n1 < 0) n1 = 0; }
						s += (int)(n1 * ge1 + (n - 1 - n1) * ext);
					}
				}
				Cu.ub[k] = s;
			}
			__syncwarp(gmask);
			store_block(Cu, head);
			cur_node = head;
		}
		uint32_t sp = 0;
		mpos[head] = -1;
		stack[sp++] = head;
		__syncwarp(gmask);

		while(sp){
			const uint32_t u = stack[--sp];
			const int2 un = node[u];
			const int rpos_u = un.x; const uint32_t base_u = ((uint32_t)un.y >> 16) & 0xffu;
			const int mpos_u = mpos[u];
			const int e0 = eoff[u], e1 = eoff[u + 1];
			// the predecessor row: the row just computed (backbone case) or a block read back from HBM
			__syncwarp(gmask);
			if(cur_node == u){ PoaSlot t = Pd; Pd = Cu; Cu = t; cur_node = 0xffffffffu; }
			else load_block(Pd, u);
			__syncwarp(gmask);
			for(int ei=e0;ei<e1;ei++){
				const uint32_t v = (uint32_t)edst[ei];
				int mpos_v = mpos[v];
				if(mpos_u + 1 < mpos_v){ mpos_v = mpos_u + 1; mpos[v] = mpos_v; }
				if(v == tail){
					// end candidates (bspoa.h:2548-2580)
					int moff = (int)min((int64_t)slen, (int64_t)rpos_u + (int64_t)bw) - 1;
					int smx = getscore(Pd, (int64_t)moff - rpos_u);
					if((int)slen > moff + 1){
						int rem = (int)slen - moff - 1;
						if(pw < 2) smx += O + E * rem;
						else smx += max(O + E * rem, Q + P * rem);
					}
					smx += T;
					if(smx > maxscr){ maxscr = smx; maxidx = (int)u; maxoff = moff; }
					if(mode == 1){
						// row_max with the SSE reduction's tie-break order (bsalign.h:3213-3291)
						const uint32_t nck = (W + 31) / 32;
						int Max = kScoreMin, Scr = Pd.ub[j]; uint32_t Idx = (uint32_t)j;
						for(uint32_t c=0;c<nck;c++){
							uint32_t lo = c * 32, hi = lo + 32 < W ? lo + 32 : W;
							int run = 0, mx = -32767;
							for(uint32_t i=lo;i<hi;i++){ run += Pd.u[i * 16 + j]; if(run > mx) mx = run; }
							int hh = Scr + mx;
							if(hh > Max){ Max = hh; Idx = (uint32_t)j | (c << 8); }
							Scr += run;
						}
						sRM[j] = Max; sRM[16 + j] = (int)Idx;
						__syncwarp(gmask);
						int M4[4]; uint32_t I4[4];
						#pragma unroll
						for(int k=0;k<4;k++){
							int m0 = sRM[k], m1 = sRM[k + 8];
							uint32_t i0 = (uint32_t)sRM[16 + k], i1 = (uint32_t)sRM[16 + k + 8];
							if(sRM[k + 4] > m0){ m0 = sRM[k + 4]; i0 = (uint32_t)sRM[16 + k + 4]; }
							if(sRM[k + 12] > m1){ m1 = sRM[k + 12]; i1 = (uint32_t)sRM[16 + k + 12]; }
							if(m1 > m0){ m0 = m1; i0 = i1; }
							M4[k] = m0; I4[k] = i0;
						}
						int max_score = M4[0]; uint32_t bi = I4[0];
						#pragma unroll
						for(int k=1;k<4;k++) if(M4[k] > max_score){ max_score = M4[k]; bi = I4[k]; }
						__syncwarp(gmask);
						if(max_score > maxscr){
							uint32_t bl = bi & 0xff, bc = bi >> 8;
							uint32_t x = bc * 32, y = (bc + 1) * 32 < W ? (bc + 1) * 32 : W;
							uint32_t pos = x; int umax = kScoreMin, uscr = 0;
							for(;x<y;x++){ uscr += Pd.u[x * 16 + bl]; if(uscr > umax){ pos = x; umax = uscr; } }
							maxscr = max_score; maxidx = (int)u; maxoff = (int)(bl * W + pos) + rpos_u;
						}
					}
					vst[v] = vst[v] + 1;
					continue;
				}
				// ================= dpalign_row_update_bspoa (bspoa.h:2232-2261) =================================
				const int2 vn = node[v];
				const int rpos_v = vn.x;
				const uint32_t nct_v = (uint32_t)vn.y & 0xffffu, base_v = ((uint32_t)vn.y >> 16) & 0xffu, bonus_v = ((uint32_t)vn.y >> 24) & 1u;
				const uint32_t vst_v = vst[v];
				const uint32_t m = (uint32_t)rpos_v - (uint32_t)rpos_u;
				const uint32_t q1 = (uint32_t)rpos_u, q2 = (uint32_t)rpos_v;
				// substitution scores: profile (v.base == u.base) * 2 + v.bonus (bspoa.h:2588, 2199-2215)
				const int Mk = (int8_t)(Mm + (bonus_v ? refbonus : 0)), Xk = (int8_t)Xx;
				const uint32_t hpc = (base_v != base_u) ? 1u : 0u;
				// ---- shifted anchors (bsalign.h:2253-2269, 2310-2331, 2350-2389) ----
				const uint32_t cyc = (m < bw) ? m / W : 0, mr = (m < bw) ? m - cyc * W : 0;
				const uint32_t i0 = bw - m;   // first band position the old row does not cover (m < bw)
				for(int k=j;k<=kLanes;k+=kPoaGroup){
					int s;
					if(m >= bw) s = kScoreMin;
					else if(m == 0) s = Pd.ub[k];
					else {
						if(k + cyc < (uint32_t)kLanes){
							s = Pd.ub[k + cyc];
							for(uint32_t t=0;t<mr;t++) s += Pd.u[t * 16 + k + cyc];
						} else s = Pd.ub[kLanes];
						uint32_t Pp = (uint32_t)k * W;
						if(k >= 1 && Pp > i0){
							uint32_t kk = Pp - i0;
							uint32_t n1 = (kk - 1 < ovd - 1) ? kk - 1 : ovd - 1;
							s += ovc + (int)n1 * ge1 + (int)(kk - 1 - n1) * ge2;
						}
					}
					sUBS[k] = s;
				}
				__syncwarp(gmask);
				int rh;
				if(q1 == q2){
					if(q1) rh = kScoreMin;
					else if(mode == 1 || mpos_v == 0) rh = 0;
					else if(pw < 2) rh = O + E * mpos_v;
					else rh = max(O + E * mpos_v, Q + P * mpos_v);
				} else if(q1 + bw >= q2) rh = sUBS[0];
				else rh = kScoreMin;
				// previous-row cell of this lane's step i after the shift: own lane / right neighbours / overhang / zeros
				#define POA_PREV(i, pu, pe, pq) { \
					uint32_t pp_ = (uint32_t)j * W + (i) + m; \
					if(m >= bw){ pu = 0; pe = 0; pq = 0; } \
					else if(pp_ < bw){ uint32_t jj_ = pp_ / W, ii_ = pp_ - jj_ * W; pu = Pd.u[ii_ * 16 + jj_]; pe = Pd.e[ii_ * 16 + jj_]; pq = Pd.q[ii_ * 16 + jj_]; } \
					else { uint32_t k_ = pp_ - bw; pu = (int8_t)(k_ == 0 ? ovc : (k_ < ovd ? ge1 : ge2)); pe = 0; pq = 0; } }
				#define POA_Z(i, z) { \
					uint32_t x_ = q2 + (uint32_t)j * W + (i); \
					if(x_ < slen){ uint32_t c_ = qc[x_]; z = (int8_t)(((c_ & 3u) == base_v ? Mk : Xk) + (int)(hpc & (c_ >> 2))); } \
					else z = kEpi8Min; }
				// ---- cell 0 of the band (bsalign.h:2899-2907) ----
				int h0;
				{
					int pu, pe, pq, z0, t0;
					uint32_t pp0 = m;
					if(m >= bw){ pu = 0; pe = 0; pq = 0; }
					else { uint32_t jj_ = pp0 / W, ii_ = pp0 - jj_ * W; pu = Pd.u[ii_ * 16 + jj_]; pe = Pd.e[ii_ * 16 + jj_]; pq = Pd.q[ii_ * 16 + jj_]; }
					if(q2 < slen){ uint32_t c_ = qc[q2]; z0 = (int8_t)(((c_ & 3u) == base_v ? Mk : Xk) + (int)(hpc & (c_ >> 2))); } else z0 = kEpi8Min;
					h0 = (rh - sUBS[0]) + z0;
					if(pw == 0) t0 = pu + ge1;
					else if(pw == 1) t0 = pu + pe;
					else t0 = pu + max(pe, pq);
					if(h0 >= t0){ if(h0 > kEpi8Max) h0 = kEpi8Max; } else h0 = kEpi8Min;
				}
				// ---- pass 1: F (G) leaving the lane's running block with nothing entering ----
				{
					int f = kEpi8Min, g = kEpi8Min;
					for(uint32_t i=0;i<W;i++){
						int pu, pe, pq, z, h;
						POA_PREV(i, pu, pe, pq)
						POA_Z(i, z)
						if(j == 0 && i == 0) z = h0;
						if(pw == 0){
							int e = clamp8(pu + ge1);
							h = max(max(e, z), f);
							f = clamp8(clamp8(h + ge1) - pu);
						} else if(pw == 1){
							int e = clamp8(pe + pu);
							h = max(max(e, z), f);
							f = clamp8(max(clamp8(f + ge1), clamp8(h + GOE)) - pu);
						} else {
							int e = clamp8(pe + pu), q = clamp8(pq + pu);
							h = max(max(max(e, z), max(q, f)), g);
							int hh = clamp8(h + GOE);
							f = clamp8(max(clamp8(f + ge1), hh) - pu);
							hh = clamp8(hh - GOQ);
							g = clamp8(max(clamp8(g + ge2), hh) - pu);
						}
					}
					sF[j] = (int8_t)f; sF[16 + j] = (int8_t)g;
				}
				__syncwarp(gmask);
				// ---- F penetration (bsalign.h:2639-2652): exact 16-step scan, every thread redundantly ----
				int fin = kEpi8Min, gin = kEpi8Min;
				{
					const int tW = (int)W * ge1, tW2 = (int)W * ge2;
					int ubp = sUBS[0], ubn = sUBS[1];
					int s = tW + kEpi8Min - (ubn - ubp), s2 = tW2 + kEpi8Min - (ubn - ubp);
					for(int k=1;k<kLanes;k++){
						int fk = sF[k - 1];
						if(fk < s) fk = (int)(int8_t)s;
						int gk = 0;
						if(pw == 2){ gk = sF[16 + k - 1]; if(gk < s2) gk = (int)(int8_t)s2; }
						if(k == j){ fin = fk; gin = gk; }
						ubp = ubn; ubn = sUBS[k + 1];
						s = tW + fk - (ubn - ubp);
						if(pw == 2) s2 = tW2 + gk - (ubn - ubp);
					}
				}
				// ---- pass 2: the row (bsalign.h:2934-2957, 3141-3176) into the current slot ----
				__syncwarp(gmask);   // every lane has read what it needs from Cu's previous content? (Cu is never a source here)
				int vt, un0 = 0;
				{
					int f = fin, g = (pw == 2) ? gin : 0, h = 0, pu = 0, vv = 0;
					for(uint32_t i=0;i<W;i++){
						int pe, pq, z, e, q, unew;
						POA_PREV(i, pu, pe, pq)
						POA_Z(i, z)
						if(j == 0 && i == 0) z = h0;
						if(pw == 0){
							e = clamp8(pu + ge1);
							h = max(max(e, z), f);
							unew = clamp8(h - vv);
							vv = clamp8(h - pu);
							f = clamp8(clamp8(h + ge1) - pu);
						} else if(pw == 1){
							e = clamp8(pe + pu);
							h = max(max(e, z), f);
							unew = clamp8(h - vv);
							vv = clamp8(h - pu);
							Cu.e[i * 16 + j] = (int8_t)max(clamp8(clamp8(e + ge1) - h), GOE);
							h = clamp8(h + GOE);
							f = clamp8(max(clamp8(f + ge1), h) - pu);
						} else {
							e = clamp8(pe + pu); q = clamp8(pq + pu);
							h = max(max(max(e, z), max(q, f)), g);
							unew = clamp8(h - vv);
							vv = clamp8(h - pu);
							Cu.e[i * 16 + j] = (int8_t)max(clamp8(clamp8(e + ge1) - h), GOE);
							Cu.q[i * 16 + j] = (int8_t)max(clamp8(clamp8(q + ge2) - h), GQP);
							h = clamp8(h + GOE);
							f = clamp8(max(clamp8(f + ge1), h) - pu);
							h = clamp8(h - GOQ);
							g = clamp8(max(clamp8(g + ge2), h) - pu);
						}
						if(i == 0) un0 = unew; else Cu.u[i * 16 + j] = (int8_t)unew;
					}
					// the SSE code leaves h biased by the gap-open constant after the loop and removes it (bsalign.h:2958, 3177)
					if(pw == 1) h = clamp8(h - GOE);
					else if(pw == 2) h = clamp8(h - GQP);
					vt = clamp8(h - pu);
				}
				#undef POA_PREV
				#undef POA_Z
				// ---- tail (bsalign.h:2618-2636) ----
				{
					int vprev = __shfl_up_sync(gmask, vt, 1, kPoaGroup);
					if(j == 0) vprev = 0;
					int u0n = clamp8(un0 - vprev);
					Cu.ub[j + 1] = sUBS[j + 1] + vt;
					if(j == 0){ Cu.ub[0] = sUBS[0] + u0n; u0n = 0; }
					Cu.u[j] = (int8_t)u0n;
				}
				nupd++;
				__syncwarp(gmask);
				// ---- first visit: the block of v; later visits: merge into it (bspoa.h:2263-2272, bsalign.h:2474-2616) ----
				if(vst_v){
					const uint8_t *vb = rows + (size_t)v * mmblk;
					const int8_t *bu = (const int8_t*)vb, *be = bu + bw, *bq = be + bw;
					const int32_t *bub = (const int32_t*)(vb + bw * (pw + 1));
					int sa = Cu.ub[j], sb = bub[j];
					const int ub16a = Cu.ub[16], ub16b = bub[16];
					for(uint32_t ib=0;ib<W;){
						const uint32_t ie = min(ib + 256u, W);
						int d = max(-0x7FFF, min(0x7FFF, sa - sb));
						int xa = d >> 1, xb = xa - d;
						sa -= xa; sb -= xb;
						int ta = xa, tb = xb, mp = max(ta, tb);
						for(uint32_t i=ib;i<ie;i++){
							const uint32_t o = i * 16 + j;
							ta = sat16i(ta + Cu.u[o]); tb = sat16i(tb + bu[o]);
							const int mc = max(ta, tb);
							if(pw >= 1){ int ya = sat16i(ta + Cu.e[o]), yb = sat16i(tb + be[o]); Cu.e[o] = (int8_t)clamp8(sat16i(max(ya, yb) - mc)); }
							if(pw == 2){ int ya = sat16i(ta + Cu.q[o]), yb = sat16i(tb + bq[o]); Cu.q[o] = (int8_t)clamp8(sat16i(max(ya, yb) - mc)); }
							Cu.u[o] = (int8_t)clamp8(sat16i(mc - mp));
							mp = mc;
						}
						sa += ta; sb += tb;
						ib = ie;
					}
					__syncwarp(gmask);
					Cu.ub[j] = max(Cu.ub[j], bub[j]);
					if(j == 0) Cu.ub[16] = max(ub16a, ub16b);
					nmrg++;
					__syncwarp(gmask);
				}
				store_block(Cu, v);
				cur_node = v;
				vst[v] = vst_v + 1;
				if(vst_v + 1 == nct_v){
					if(mode != 0 && q2 + bw >= slen){
						int smx = getscore(Cu, (int64_t)slen - 1 - (int64_t)q2) + T;
						if(smx > maxscr){ maxscr = smx; maxidx = (int)v; maxoff = (int)slen - 1; }
					}
					stack[sp++] = v;
				}
				__syncwarp(gmask);
			}
		}
		if(j == 0){
			a.best[(size_t)job * 3 + 0] = maxscr; a.best[(size_t)job * 3 + 1] = maxidx; a.best[(size_t)job * 3 + 2] = maxoff;
			a.status[job] = stflag;
			a.ops[(size_t)job * 2 + 0] = nupd; a.ops[(size_t)job * 2 + 1] = nmrg;
		}
		__syncwarp(gmask);
	}
}

} // namespace bsb200
