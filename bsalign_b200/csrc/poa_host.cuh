// poa_host.cuh -- host side of the POA sweep entry points (included by bsb200.cu after the context definition).
// Uploads a batch of sweep jobs (CSR sub-graphs + reads), runs poa_prep_kernel + poa_sweep_kernel on the context's
// stream and brings back the node row blocks (reference memp layout), the best end per job and the op counts.
#pragma once
#include "poa_kernels.cuh"
#include "poa_fast.cuh"
#include "poa_backtrace.cuh"

struct bsb200_poa_batch {
	uint32_t njobs = 0;
	uint64_t nnodes = 0, nedges = 0, qbytes = 0, row_bytes = 0;
	uint32_t max_bw = 16;
	std::vector<uint64_t> row_off;      // njobs + 1
	std::vector<uint32_t> order;        // generic-kernel jobs, heaviest first
	std::vector<uint32_t> order_fast[3]; // register-kernel jobs by piecewise (1, 2), heaviest first
	std::vector<uint64_t> qsel_off;
	uint64_t qsel_bytes = 0;
	DevBuf d_order_fast, d_qsel, d_qsel_off, d_counter2;
	HostBuf h_nodes;
	bool has_rev = false;               // reverse edges attached: the run also walks the alignment back on the device
	uint64_t nredges = 0;
	DevBuf d_reoff, d_redge_off, d_resrc, d_recov, d_rev, d_match, d_trace;
	DevBuf d_par, d_queries, d_qcode, d_qoff, d_slen, d_node_off, d_node, d_eoff, d_edge_off, d_edst, d_head, d_tail,
		d_mpos, d_vst, d_stack, d_rows, d_row_off, d_best, d_status, d_ops, d_order, d_counter;
	bool ran = false;
};

static std::vector<DevBuf*> poa_bufs(bsb200_poa_batch *b){
	return {&b->d_par, &b->d_queries, &b->d_qcode, &b->d_qoff, &b->d_slen, &b->d_node_off, &b->d_node, &b->d_eoff, &b->d_edge_off, &b->d_edst,
		&b->d_head, &b->d_tail, &b->d_mpos, &b->d_vst, &b->d_stack, &b->d_rows, &b->d_row_off, &b->d_best, &b->d_status, &b->d_ops, &b->d_order, &b->d_counter,
		&b->d_order_fast, &b->d_qsel, &b->d_qsel_off, &b->d_counter2, &b->d_reoff, &b->d_redge_off, &b->d_resrc, &b->d_recov, &b->d_rev, &b->d_match, &b->d_trace};
}

static uint32_t poa_mmblk(const int32_t *par){
	const uint32_t bw = (uint32_t)par[0];
	const int pw = epi8_piecewise((int8_t)par[4], (int8_t)par[5], (int8_t)par[6], (int8_t)par[7], (int)bw);
	return (bw * (pw + 1) + 68 + 15) / 16 * 16;   // bspoa.h:2217
}

extern "C" uint32_t bsb200_poa_block_bytes(const int32_t par[10]){ return par ? poa_mmblk(par) : 0; }

extern "C" void bsb200_poa_free(bsb200_ctx *ctx, bsb200_poa_batch *b){
	if(!b) return;
	if(ctx) cudaSetDevice(ctx->device);
	std::vector<DevBuf*> ds = poa_bufs(b);
	if(ctx){   // park the allocations for the next batch (cudaMalloc / cudaFree of a multi-GB row arena cost more than the sweep)
		cudaStreamSynchronize(ctx->stream);
		for(size_t k=0;k<ds.size();k++){ if(ctx->poa_cache[k].cap < ds[k]->cap){ ctx->poa_cache[k].release(); ctx->poa_cache[k] = *ds[k]; } else ds[k]->release(); }
		if(ctx->poa_hcache[0].cap < b->h_nodes.cap){ ctx->poa_hcache[0].release(); ctx->poa_hcache[0] = b->h_nodes; } else b->h_nodes.release();
	} else {
		for(auto d : ds) d->release();
		b->h_nodes.release();
	}
	delete b;
}

extern "C" bsb200_poa_batch *bsb200_poa_upload(bsb200_ctx *ctx, uint32_t njobs, const int32_t *par,
		const uint8_t *queries, const uint64_t *qoff, const uint32_t *slen,
		const uint64_t *node_off, const uint8_t *node_base, const uint8_t *node_bonus, const int32_t *node_rpos, const int32_t *node_nct,
		const int32_t *eoff, const uint64_t *edge_off, const int32_t *edst, const uint32_t *head, const uint32_t *tail){
	if(!ctx) return nullptr;
	ctx->err.clear();
	if(njobs && (!par || !queries || !qoff || !slen || !node_off || !node_base || !node_bonus || !node_rpos || !node_nct || !eoff || !edge_off || !edst || !head || !tail)){
		fail(ctx, "bsb200_poa_upload", cudaSuccess); return nullptr;
	}
	cudaSetDevice(ctx->device);
	bsb200_poa_batch *b = new bsb200_poa_batch();
	{
		std::vector<DevBuf*> ds = poa_bufs(b);
		for(size_t k=0;k<ds.size();k++){ *ds[k] = ctx->poa_cache[k]; ctx->poa_cache[k] = DevBuf(); }
		b->h_nodes = ctx->poa_hcache[0]; ctx->poa_hcache[0] = HostBuf();
	}
	b->njobs = njobs;
	b->row_off.assign((size_t)njobs + 1, 0);
	b->nnodes = njobs ? node_off[njobs] : 0;
	b->nedges = njobs ? edge_off[njobs] : 0;
	std::vector<std::pair<uint64_t, uint32_t>> kv(njobs);
	for(uint32_t i=0;i<njobs;i++){
		const int32_t *p = par + (size_t)i * 10;
		const uint64_t nn = node_off[i + 1] - node_off[i];
		if(p[0] <= 0 || p[0] % 16 || nn == 0 || head[i] >= nn || tail[i] >= nn || nn > 0x7fffffffull){
			ctx->err = "bsb200_poa_upload: bad job (bandwidth must be a positive multiple of 16; head/tail must be local node ids)";
			bsb200_poa_free(ctx, b); return nullptr;
		}
		b->max_bw = std::max<uint32_t>(b->max_bw, (uint32_t)p[0]);
		b->row_off[i + 1] = b->row_off[i] + nn * poa_mmblk(p);
		b->qbytes = std::max<uint64_t>(b->qbytes, qoff[i] + slen[i]);
		kv[i] = std::make_pair(~(nn * (uint64_t)p[0]), i);
	}
	b->row_bytes = b->row_off[njobs];
	std::sort(kv.begin(), kv.end());
	// node records (rpos, nct | base << 16 | bonus << 24) into pinned staging, and which jobs the register-resident kernel takes
	// (poa_fast.cuh): band of exactly 128 cells that never runs past the read end, gap costs <= 0 with a gap-open cost, small scores;
	// everything else goes to the generic kernel.  Jobs are independent: a few host threads share them.
	if(b->h_nodes.reserve(b->nnodes * 8 + 16) != cudaSuccess){ fail(ctx, "pinned allocation (poa)", cudaErrorMemoryAllocation); bsb200_poa_free(ctx, b); return nullptr; }
	int2 *nodes = b->h_nodes.as<int2>();
	b->qsel_off.assign((size_t)njobs + 1, 0);
	std::vector<uint8_t> fast_pw(njobs, 0);
	{
		std::atomic<uint32_t> next(0);
		auto work = [&](){
			while(true){
				const uint32_t i = next.fetch_add(1);
				if(i >= njobs) break;
				const int32_t *p = par + (size_t)i * 10;
				const int O = p[4], E = p[5], Q = p[6], P = p[7];
				const int pw = epi8_piecewise((int8_t)O, (int8_t)E, (int8_t)Q, (int8_t)P, p[0]);
				bool ok = p[0] == kPoaFastBw && slen[i] >= (uint32_t)kPoaFastBw && pw >= 1 && O >= -64 && O <= 0 && E >= -64 && E <= 0 && O + E <= 0 &&
					p[2] + p[9] + 1 <= 62 && p[2] >= -62 && p[3] >= -62 && p[3] <= 62 && p[9] >= 0;
				if(pw == 2) ok = ok && Q >= -64 && Q <= 0 && P >= -64 && P <= 0;
				const uint64_t k0 = node_off[i], k1 = node_off[i + 1];
				const int64_t lim = (int64_t)slen[i] - kPoaFastBw;
				bool in_band = true, bases = true;
				for(uint64_t k=k0;k<k1;k++){
					nodes[k].x = node_rpos[k];
					nodes[k].y = (int)(((uint32_t)node_nct[k] & 0xffffu) | ((uint32_t)node_base[k] << 16) | (((uint32_t)node_bonus[k] & 1u) << 24));
					in_band &= (node_rpos[k] >= 0) & ((int64_t)node_rpos[k] <= lim);
					bases &= (node_base[k] <= 3) | (k - k0 == tail[i]) | (k - k0 == head[i]);
				}
				if(ok && in_band && bases) fast_pw[i] = (uint8_t)pw;
			}
		};
		const uint32_t nth = std::max(1u, std::min(16u, std::min(njobs, std::thread::hardware_concurrency())));
		std::vector<std::thread> th;
		for(uint32_t k=1;k<nth;k++) th.emplace_back(work);
		work();
		for(auto &x : th) x.join();
	}
	for(uint32_t i=0;i<njobs;i++) b->qsel_off[i + 1] = b->qsel_off[i] + (fast_pw[i] ? ((uint64_t)slen[i] + 16 + 15) / 16 * 16 : 0);
	b->qsel_bytes = b->qsel_off[njobs];
	for(uint32_t i=0;i<njobs;i++){
		const uint32_t jb = kv[i].second;
		if(fast_pw[jb]) b->order_fast[fast_pw[jb]].push_back(jb); else b->order.push_back(jb);
	}
	cudaError_t e = cudaSuccess;
	auto R = [&](cudaError_t x){ if(e == cudaSuccess) e = x; };
	const size_t nj = njobs, nn = b->nnodes, ne = b->nedges;
	R(b->d_par.reserve(nj * 40 + 16)); R(b->d_queries.reserve(b->qbytes + 16)); R(b->d_qcode.reserve(b->qbytes + 16));
	R(b->d_qoff.reserve(nj * 8 + 8)); R(b->d_slen.reserve(nj * 4 + 4)); R(b->d_node_off.reserve((nj + 1) * 8)); R(b->d_node.reserve(nn * 8 + 8));
	R(b->d_eoff.reserve((nn + nj) * 4 + 4)); R(b->d_edge_off.reserve((nj + 1) * 8)); R(b->d_edst.reserve(ne * 4 + 4));
	R(b->d_head.reserve(nj * 4 + 4)); R(b->d_tail.reserve(nj * 4 + 4));
	R(b->d_mpos.reserve(nn * 4 + 4)); R(b->d_vst.reserve(nn * 4 + 4)); R(b->d_stack.reserve(nn * 4 + 4));
	R(b->d_rows.reserve(b->row_bytes + 16)); R(b->d_row_off.reserve((nj + 1) * 8));
	R(b->d_best.reserve(nj * 12 + 12)); R(b->d_status.reserve(nj * 4 + 4)); R(b->d_ops.reserve(nj * 16 + 16));
	R(b->d_order.reserve(nj * 4 + 4)); R(b->d_counter.reserve(256)); R(b->d_counter2.reserve(256));
	R(b->d_order_fast.reserve(nj * 4 + 4)); R(b->d_qsel.reserve(b->qsel_bytes + 64)); R(b->d_qsel_off.reserve((nj + 1) * 8));
	if(e != cudaSuccess){ fail(ctx, "device allocation (poa)", e); bsb200_poa_free(ctx, b); return nullptr; }
	cudaStream_t st = ctx->stream;
	cudaEventRecord(ctx->ev[0], st);
	if(njobs){
		R(cudaMemcpyAsync(b->d_par.p, par, nj * 40, cudaMemcpyHostToDevice, st));
		R(cudaMemcpyAsync(b->d_queries.p, queries, b->qbytes, cudaMemcpyHostToDevice, st));
		R(cudaMemcpyAsync(b->d_qoff.p, qoff, nj * 8, cudaMemcpyHostToDevice, st));
		R(cudaMemcpyAsync(b->d_slen.p, slen, nj * 4, cudaMemcpyHostToDevice, st));
		R(cudaMemcpyAsync(b->d_node_off.p, node_off, (nj + 1) * 8, cudaMemcpyHostToDevice, st));
		R(cudaMemcpyAsync(b->d_node.p, nodes, nn * 8, cudaMemcpyHostToDevice, st));
		R(cudaMemcpyAsync(b->d_eoff.p, eoff, (nn + nj) * 4, cudaMemcpyHostToDevice, st));
		R(cudaMemcpyAsync(b->d_edge_off.p, edge_off, (nj + 1) * 8, cudaMemcpyHostToDevice, st));
		if(ne) R(cudaMemcpyAsync(b->d_edst.p, edst, ne * 4, cudaMemcpyHostToDevice, st));
		R(cudaMemcpyAsync(b->d_head.p, head, nj * 4, cudaMemcpyHostToDevice, st));
		R(cudaMemcpyAsync(b->d_tail.p, tail, nj * 4, cudaMemcpyHostToDevice, st));
		R(cudaMemcpyAsync(b->d_row_off.p, b->row_off.data(), (nj + 1) * 8, cudaMemcpyHostToDevice, st));
		if(!b->order.empty()) R(cudaMemcpyAsync(b->d_order.p, b->order.data(), b->order.size() * 4, cudaMemcpyHostToDevice, st));
		{
			size_t at = 0;
			for(int pw=1;pw<=2;pw++){
				if(!b->order_fast[pw].empty()) R(cudaMemcpyAsync(b->d_order_fast.as<uint32_t>() + at, b->order_fast[pw].data(), b->order_fast[pw].size() * 4, cudaMemcpyHostToDevice, st));
				at += b->order_fast[pw].size();
			}
		}
		R(cudaMemcpyAsync(b->d_qsel_off.p, b->qsel_off.data(), (nj + 1) * 8, cudaMemcpyHostToDevice, st));
	}
	cudaEventRecord(ctx->ev[1], st);
	R(cudaStreamSynchronize(st));
	if(e != cudaSuccess){ fail(ctx, "host to device copy (poa)", e); bsb200_poa_free(ctx, b); return nullptr; }
	float ms = 0; cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[1]);
	ctx->timing = bsb200_timing_t();
	ctx->timing.h2d_ms = ms;
	ctx->timing.h2d_bytes = nj * (40 + 8 + 4 + 8 + 8 + 4 + 4 + 8 + 4) + b->qbytes + nn * 8 + (nn + nj) * 4 + ne * 4;
	return b;
}

extern "C" int bsb200_poa_run(bsb200_ctx *ctx, bsb200_poa_batch *b){
	if(!ctx || !b) return -1;
	ctx->err.clear();
	cudaSetDevice(ctx->device);
	cudaStream_t st = ctx->stream;
	if(b->njobs == 0){ b->ran = true; return 0; }
	PoaArgs a;
	a.njobs = b->njobs; a.order = b->d_order.as<uint32_t>(); a.counter = b->d_counter.as<unsigned int>();
	a.par = b->d_par.as<int32_t>(); a.qcode = b->d_qcode.as<uint8_t>(); a.qoff = b->d_qoff.as<uint64_t>(); a.slen = b->d_slen.as<uint32_t>();
	a.node_off = b->d_node_off.as<uint64_t>(); a.node = b->d_node.as<int2>(); a.eoff = b->d_eoff.as<int32_t>();
	a.edge_off = b->d_edge_off.as<uint64_t>(); a.edst = b->d_edst.as<int32_t>(); a.head = b->d_head.as<uint32_t>(); a.tail = b->d_tail.as<uint32_t>();
	a.mpos = b->d_mpos.as<int32_t>(); a.vst = b->d_vst.as<uint32_t>(); a.stack = b->d_stack.as<uint32_t>();
	a.rows = b->d_rows.as<uint8_t>(); a.row_off = b->d_row_off.as<uint64_t>();
	a.best = b->d_best.as<int32_t>(); a.status = b->d_status.as<int32_t>(); a.ops = b->d_ops.as<unsigned long long>();
	a.slot_bytes = 3 * b->max_bw + 80;
	const size_t smem = (size_t)(kPoaThreads / kPoaGroup) * (2 * a.slot_bytes + 80 + 32 + 128);
	const uint32_t ngen = (uint32_t)b->order.size();
	if(ngen && smem > ctx->smem_optin){ ctx->err = "bsb200_poa_run: bandwidth too wide for the shared-memory row slots"; return -1; }
	if(ngen) CK(cudaFuncSetAttribute(poa_sweep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
	cudaEventRecord(ctx->ev[2], st);
	CK(cudaMemsetAsync(b->d_counter.p, 0, 16, st));
	CK(cudaMemsetAsync(b->d_counter2.p, 0, 16, st));
	uint32_t launches = 0;
	if(ngen){
		const uint32_t prep_blocks = std::min<uint32_t>(ngen, (uint32_t)ctx->num_sms * 8);
		a.njobs = ngen;
		poa_prep_kernel<<<prep_blocks, 256, 0, st>>>(ngen, a.order, b->d_queries.as<uint8_t>(), a.qoff, a.slen, b->d_qcode.as<uint8_t>(), a.node_off, a.mpos, a.vst);
	}
	const uint32_t nfast = (uint32_t)(b->order_fast[1].size() + b->order_fast[2].size());
	if(nfast){
		const uint32_t prep_blocks = std::min<uint32_t>(nfast, (uint32_t)ctx->num_sms * 8);
		poa_fast_prep_kernel<<<prep_blocks, 256, 0, st>>>(nfast, b->d_order_fast.as<uint32_t>(), b->d_queries.as<uint8_t>(), a.qoff, a.slen,
			b->d_qsel.as<uint8_t>(), b->d_qsel_off.as<uint64_t>(), a.node_off, a.mpos, a.vst);
	}
	cudaEventRecord(ctx->ev[3], st);
	if(ngen){
		const uint32_t groups = kPoaThreads / kPoaGroup;
		uint32_t blocks = (ngen + groups - 1) / groups;
		blocks = std::min<uint32_t>(blocks, (uint32_t)ctx->num_sms * 16);
		a.njobs = ngen;
		poa_sweep_kernel<<<blocks, kPoaThreads, smem, st>>>(a);
		launches++;
	}
	{
		size_t at = 0;
		for(int pw=1;pw<=2;pw++){
			const uint32_t nf = (uint32_t)b->order_fast[pw].size();
			if(!nf) continue;
			PoaFastArgs fa;
			fa.base = a; fa.base.njobs = nf; fa.base.order = b->d_order_fast.as<uint32_t>() + at;
			fa.base.counter = b->d_counter2.as<unsigned int>() + (pw - 1);
			fa.qsel = b->d_qsel.as<uint32_t>(); fa.qsel_off = b->d_qsel_off.as<uint64_t>(); fa.all_ones = 0xffffffffu;
			// jobs per warp: one while every job can have a resident warp of its own (the sweep is a latency-bound chain), four when
			// the batch is large enough to be issue bound (BSB200_POA_GPW overrides, for experiments)
			int gpw = nf > (uint32_t)ctx->num_sms * 12 ? 4 : 1;
			if(const char *ev = getenv("BSB200_POA_GPW")){ int g = atoi(ev); if(g == 1 || g == 2 || g == 4) gpw = g; }
			const uint32_t blocks = std::min<uint32_t>((nf + gpw - 1) / gpw, (uint32_t)ctx->num_sms * 32);
			const size_t fsm = (size_t)kPoaFastSmem * gpw;
			#define POA_FAST_LAUNCH(PWV) do { \
				if(gpw == 4) poa_sweep_fast_kernel<PWV, 4><<<blocks, 32, fsm, st>>>(fa); \
				else if(gpw == 2) poa_sweep_fast_kernel<PWV, 2><<<blocks, 32, fsm, st>>>(fa); \
				else poa_sweep_fast_kernel<PWV, 1><<<blocks, 32, fsm, st>>>(fa); } while(0)
			if(pw == 1) POA_FAST_LAUNCH(1); else POA_FAST_LAUNCH(2);
			#undef POA_FAST_LAUNCH
			launches++;
			at += nf;
		}
	}
	cudaEventRecord(ctx->ev[4], st);
	if(b->has_rev){
		PoaBtArgs t;
		t.njobs = b->njobs; t.par = a.par; t.queries = b->d_queries.as<uint8_t>(); t.qoff = a.qoff; t.slen = a.slen;
		t.node_off = a.node_off; t.node = a.node; t.reoff = b->d_reoff.as<int32_t>(); t.redge_off = b->d_redge_off.as<uint64_t>();
		t.rev = b->d_rev.as<int4>(); t.head = a.head; t.tail = a.tail; t.rows = a.rows; t.row_off = a.row_off; t.best = a.best;
		t.match = b->d_match.as<int32_t>(); t.trace = b->d_trace.as<int32_t>();
		if(b->max_bw == 128 && b->order.empty()) poa_backtrace_kernel<true><<<b->njobs, 32, 0, st>>>(t);   // every job has the default band
		else poa_backtrace_kernel<false><<<b->njobs, 32, 0, st>>>(t);
		launches++;
	}
	cudaEventRecord(ctx->ev[7], st);
	CK(cudaGetLastError());
	CK(cudaStreamSynchronize(st));
	float prep = 0, sweep = 0, walk = 0;
	cudaEventElapsedTime(&prep, ctx->ev[2], ctx->ev[3]);
	cudaEventElapsedTime(&sweep, ctx->ev[3], ctx->ev[4]);
	cudaEventElapsedTime(&walk, ctx->ev[4], ctx->ev[7]);
	ctx->timing.forward_ms = sweep; ctx->timing.traceback_ms = walk; ctx->timing.run_ms = prep + sweep + walk; ctx->timing.total_ms = prep + sweep + walk;
	ctx->timing.traceback_launches = b->has_rev ? 1 : 0;
	ctx->timing.forward_launches = launches - (b->has_rev ? 1 : 0); ctx->timing.other_launches = (ngen ? 1 : 0) + (nfast ? 1 : 0); ctx->timing.waves = 1;
	b->ran = true;
	return 0;
}

/* rows: caller buffer of bsb200_poa_rows_bytes() bytes or NULL; row_off_out (njobs+1), best (njobs x 3), status (njobs), ops (njobs x 2) may be NULL */
extern "C" uint64_t bsb200_poa_rows_bytes(bsb200_poa_batch *b, uint64_t *row_off_out){
	if(!b) return 0;
	if(row_off_out) memcpy(row_off_out, b->row_off.data(), b->row_off.size() * 8);
	return b->row_bytes;
}

extern "C" int bsb200_poa_fetch(bsb200_ctx *ctx, bsb200_poa_batch *b, uint8_t *rows, int32_t *best, int32_t *status, uint64_t *ops){
	if(!ctx || !b) return -1;
	ctx->err.clear();
	if(!b->ran){ ctx->err = "bsb200_poa_fetch: batch has not been run"; return -1; }
	cudaSetDevice(ctx->device);
	cudaStream_t st = ctx->stream;
	cudaEventRecord(ctx->ev[5], st);
	const size_t nj = b->njobs;
	uint64_t bytes = 0;
	if(nj){
		if(rows && b->row_bytes){ CK(cudaMemcpyAsync(rows, b->d_rows.p, b->row_bytes, cudaMemcpyDeviceToHost, st)); bytes += b->row_bytes; }
		if(best){ CK(cudaMemcpyAsync(best, b->d_best.p, nj * 12, cudaMemcpyDeviceToHost, st)); bytes += nj * 12; }
		if(status){ CK(cudaMemcpyAsync(status, b->d_status.p, nj * 4, cudaMemcpyDeviceToHost, st)); bytes += nj * 4; }
		if(ops){ CK(cudaMemcpyAsync(ops, b->d_ops.p, nj * 16, cudaMemcpyDeviceToHost, st)); bytes += nj * 16; }
	}
	cudaEventRecord(ctx->ev[6], st);
	CK(cudaStreamSynchronize(st));
	float ms = 0; cudaEventElapsedTime(&ms, ctx->ev[5], ctx->ev[6]);
	ctx->timing.d2h_ms = ms; ctx->timing.d2h_bytes = bytes;
	return 0;
}

extern "C" int bsb200_poa_rows_batch(bsb200_ctx *ctx, uint32_t njobs, const int32_t *par,
		const uint8_t *queries, const uint64_t *qoff, const uint32_t *slen,
		const uint64_t *node_off, const uint8_t *node_base, const uint8_t *node_bonus, const int32_t *node_rpos, const int32_t *node_nct,
		const int32_t *eoff, const uint64_t *edge_off, const int32_t *edst, const uint32_t *head, const uint32_t *tail,
		uint8_t *rows, int32_t *best, int32_t *status, uint64_t *ops){
	bsb200_poa_batch *b = bsb200_poa_upload(ctx, njobs, par, queries, qoff, slen, node_off, node_base, node_bonus, node_rpos, node_nct, eoff, edge_off, edst, head, tail);
	if(!b) return -1;
	int rc = bsb200_poa_run(ctx, b);
	if(rc == 0) rc = bsb200_poa_fetch(ctx, b, rows, best, status, ops);
	bsb200_poa_free(ctx, b);
	return rc;
}

/* ---- device-side walk of alignment2graph_bspoa (poa_backtrace.cuh) --------------------------------------------------------- */
extern "C" int bsb200_poa_attach_reverse(bsb200_ctx *ctx, bsb200_poa_batch *b, const int32_t *reoff, const uint64_t *redge_off,
		const int32_t *resrc, const int32_t *recov){
	if(!ctx || !b) return -1;
	ctx->err.clear();
	if(b->njobs && (!reoff || !redge_off || !resrc || !recov)) return fail(ctx, "bsb200_poa_attach_reverse", cudaSuccess);
	cudaSetDevice(ctx->device);
	cudaStream_t st = ctx->stream;
	const size_t nj = b->njobs, nn = b->nnodes;
	b->nredges = nj ? redge_off[nj] : 0;
	const size_t ne = b->nredges;
	CK(b->d_reoff.reserve((nn + nj) * 4 + 4)); CK(b->d_redge_off.reserve((nj + 1) * 8)); CK(b->d_resrc.reserve(ne * 4 + 4)); CK(b->d_recov.reserve(ne * 4 + 4));
	CK(b->d_rev.reserve(ne * 32 + 32)); CK(b->d_match.reserve(b->qbytes * 4 + 16)); CK(b->d_trace.reserve(nj * 32 + 32));
	cudaEventRecord(ctx->ev[0], st);
	if(nj){
		CK(cudaMemcpyAsync(b->d_reoff.p, reoff, (nn + nj) * 4, cudaMemcpyHostToDevice, st));
		CK(cudaMemcpyAsync(b->d_redge_off.p, redge_off, (nj + 1) * 8, cudaMemcpyHostToDevice, st));
		if(ne){
			CK(cudaMemcpyAsync(b->d_resrc.p, resrc, ne * 4, cudaMemcpyHostToDevice, st));
			CK(cudaMemcpyAsync(b->d_recov.p, recov, ne * 4, cudaMemcpyHostToDevice, st));
		}
		poa_rev_prep_kernel<<<std::min<uint32_t>(b->njobs, (uint32_t)ctx->num_sms * 8), 256, 0, st>>>(b->njobs, b->d_node_off.as<uint64_t>(), b->d_node.as<int2>(),
			b->d_reoff.as<int32_t>(), b->d_redge_off.as<uint64_t>(), b->d_resrc.as<int32_t>(), b->d_recov.as<int32_t>(), b->d_rev.as<int4>());
	}
	cudaEventRecord(ctx->ev[1], st);
	CK(cudaGetLastError());
	CK(cudaStreamSynchronize(st));
	float ms = 0; cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[1]);
	ctx->timing.h2d_ms += ms;
	ctx->timing.h2d_bytes += (nn + nj) * 4 + (nj + 1) * 8 + ne * 8;
	b->has_rev = true;
	return 0;
}

/* match: int32 per read position, job i at match[qoff[i] .. + slen[i]); trace: njobs x 8 */
extern "C" int bsb200_poa_fetch_trace(bsb200_ctx *ctx, bsb200_poa_batch *b, int32_t *match, int32_t *trace){
	if(!ctx || !b) return -1;
	ctx->err.clear();
	if(!b->ran || !b->has_rev){ ctx->err = "bsb200_poa_fetch_trace: batch has not been run with reverse edges attached"; return -1; }
	cudaSetDevice(ctx->device);
	cudaStream_t st = ctx->stream;
	cudaEventRecord(ctx->ev[5], st);
	uint64_t bytes = 0;
	if(b->njobs){
		if(match && b->qbytes){ CK(cudaMemcpyAsync(match, b->d_match.p, b->qbytes * 4, cudaMemcpyDeviceToHost, st)); bytes += b->qbytes * 4; }
		if(trace){ CK(cudaMemcpyAsync(trace, b->d_trace.p, (size_t)b->njobs * 32, cudaMemcpyDeviceToHost, st)); bytes += (size_t)b->njobs * 32; }
	}
	cudaEventRecord(ctx->ev[6], st);
	CK(cudaStreamSynchronize(st));
	float ms = 0; cudaEventElapsedTime(&ms, ctx->ev[5], ctx->ev[6]);
	ctx->timing.d2h_ms += ms; ctx->timing.d2h_bytes += bytes;
	return 0;
}

/* one-shot: sweep + walk; the row blocks stay in HBM (rows may be NULL) */
extern "C" int bsb200_poa_align_batch(bsb200_ctx *ctx, uint32_t njobs, const int32_t *par,
		const uint8_t *queries, const uint64_t *qoff, const uint32_t *slen,
		const uint64_t *node_off, const uint8_t *node_base, const uint8_t *node_bonus, const int32_t *node_rpos, const int32_t *node_nct,
		const int32_t *eoff, const uint64_t *edge_off, const int32_t *edst, const uint32_t *head, const uint32_t *tail,
		const int32_t *reoff, const uint64_t *redge_off, const int32_t *resrc, const int32_t *recov,
		uint8_t *rows, int32_t *best, int32_t *status, uint64_t *ops, int32_t *match, int32_t *trace){
	bsb200_poa_batch *b = bsb200_poa_upload(ctx, njobs, par, queries, qoff, slen, node_off, node_base, node_bonus, node_rpos, node_nct, eoff, edge_off, edst, head, tail);
	if(!b) return -1;
	int rc = bsb200_poa_attach_reverse(ctx, b, reoff, redge_off, resrc, recov);
	if(rc == 0) rc = bsb200_poa_run(ctx, b);
	if(rc == 0){ ctx->timing.d2h_ms = 0; ctx->timing.d2h_bytes = 0; rc = bsb200_poa_fetch(ctx, b, rows, best, status, ops); }
	if(rc == 0) rc = bsb200_poa_fetch_trace(ctx, b, match, trace);
	bsb200_poa_free(ctx, b);
	return rc;
}
