// bsb200.cu -- host side of libbsalign_b200.so: the C ABI declared in include/bsalign_b200.h.
//
// Plans a batch (per-pair band width, heaviest-first work order, waves that fit the HBM trace budget),
// moves packed inputs to the device, launches the sm_100a kernels of epi8_kernels.cuh / edit_kernels.cuh
// on one stream and brings results + run-length CIGARs back.  No CPU fallback exists: every entry point
// fails with an error string when CUDA is unavailable.
#include "../../include/bsalign_b200.h"
#include "common.cuh"
#include "epi8_forward.cuh"
#include "epi8_wave.cuh"
#include "epi8_backcal.cuh"
#include "edit_kernels.cuh"
#include "kmer_edit.cuh"
#include "remsa_kernels.cuh"

#include <cub/device/device_scan.cuh>
#include <cub/iterator/transform_input_iterator.cuh>

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <numeric>
#include <string>
#include <thread>
#include <mutex>
#include <condition_variable>
#include <functional>
#include <atomic>
#include <vector>

using namespace bsb200;

#define BSB_VERSION "bsalign_b200 0.1 (sm_100a)"

struct DevBuf {
	void *p = nullptr; size_t cap = 0;
	// exact: no growth slack (the traceback arena is sized against the free memory)
	cudaError_t reserve(size_t n, bool exact = false){
		if(n <= cap) return cudaSuccess;
		if(p) cudaFree(p);
		p = nullptr; cap = 0;
		size_t want = (exact && n > (4ull << 30)) ? n + 256 : n + n / 8 + 256;   // exact: only the large arenas are sized against the free memory
		cudaError_t e = cudaMalloc(&p, want);
		if(e == cudaSuccess) cap = want;
		return e;
	}
	void release(){ if(p) cudaFree(p); p = nullptr; cap = 0; }
	template<class T> T *as() const { return (T*)p; }
};

struct HostBuf { // pinned staging
	void *p = nullptr; size_t cap = 0;
	cudaError_t reserve(size_t n){
		if(n <= cap) return cudaSuccess;
		if(p) cudaFreeHost(p);
		p = nullptr; cap = 0;
		size_t want = n + n / 8 + 256;
		cudaError_t e = cudaMallocHost(&p, want);
		if(e == cudaSuccess) cap = want;
		return e;
	}
	void release(){ if(p) cudaFreeHost(p); p = nullptr; cap = 0; }
	template<class T> T *as() const { return (T*)p; }
};

struct bsb200_ctx {
	int device = 0;
	int num_sms = 0;
	size_t smem_optin = 0;
	size_t total_mem = 0;
	cudaStream_t stream = nullptr;
	cudaStream_t stream_bt = nullptr;   // traceback kernels run here, concurrently with the next wave's forward kernel
	cudaEvent_t ev[8] = {};
	uint64_t trace_budget = 0;
	uint64_t auto_budget = 0;   // last automatic budget computed from cudaMemGetInfo
	std::string err;
	bsb200_timing_t timing = {};
	DevBuf trace;       // traceback arena, even waves (shared by all batches of this context; one batch runs at a time)
	DevBuf trace2;      // odd waves
	DevBuf counter;
	// allocations of finished batches are parked here and handed to the next batch (cudaMalloc/cudaFree and pinned
	// allocations cost more than the kernels of a small batch)
	DevBuf dev_cache[18];
	HostBuf host_cache[8];
	DevBuf poa_cache[40];   // same, for POA sweep batches (poa_host.cuh)
	HostBuf stage[2]; cudaEvent_t stage_ev[2] = {nullptr, nullptr};   // pinned staging for large copies from / to pageable caller memory
	bsb200_ctx *helper = nullptr;   // second context on the same device: the one-call entry points pipeline large edit batches over both
	DevBuf remsa_cache[8];  // re-alignment batches (bsb200_remsa_batch)
	DevBuf kmer_cache[8];   // k-mer guided edit: warp slots, gap-trace pool, counters, order list, fallback sub-batch results
	HostBuf poa_hcache[2];
};

struct Wave { uint32_t beg, end; uint64_t trace_bytes; };

struct bsb200_batch {
	int kind = 0;
	uint64_t n = 0;
	int mode = 0; uint32_t bandwidth = 0;
	int8_t mtx[16] = {}; int8_t go1 = 0, ge1 = 0, go2 = 0, ge2 = 0;
	int pw = 0; int want_cigar = 1;
	uint32_t max_bw = 16, max_q64 = 64, max_qlen = 0;
	int wave_split = 0;   // > 0: the batch takes the wavefront forward kernel with that many sub-blocks per lane (epi8_wave.cuh)
	std::vector<uint8_t> empty;          // 1: qlen or tlen 0; 2: edit pair too long for edit_kernel (long_pairs); 3: such a pair with a moving band (unsupported)
	std::vector<uint32_t> long_pairs;    // edit pairs that take edit_long_kernel, and their scratch offsets
	std::vector<uint64_t> long_off; uint64_t long_bytes = 0;
	uint64_t cells = 0, trace_bytes = 0, cig_words = 0, max_wave_bytes = 0;
	std::vector<Wave> waves;
	std::vector<uint32_t> order;
	std::vector<uint64_t> trace_off, cig_off;
	DevBuf d_seqs, d_qoff, d_toff, d_qlen, d_tlen, d_order, d_trace_off, d_results, d_status, d_ncigar, d_cig_raw, d_cig_off, d_cig_dense, d_dense_off, d_dense_total, d_block_rows, d_prefix, d_bits;
	HostBuf h_results, h_status, h_ncigar, h_dense_off, h_dense, h_total;
	size_t seq_bytes = 0;
	const uint8_t *d_seqs_ext = nullptr;   // device-resident arena owned by the caller (bsb200_batch_upload_dev)
	const uint8_t *seqs_dev() const { return d_seqs_ext ? d_seqs_ext : d_seqs.as<uint8_t>(); }
	bool ran = false;
};

static int fail(bsb200_ctx *ctx, const char *what, cudaError_t e){
	char buf[512];
	snprintf(buf, sizeof(buf), "%s: %s", what, e == cudaSuccess ? "invalid argument" : cudaGetErrorString(e));
	if(ctx) ctx->err = buf;
	return -1;
}
#define CK(call) do { cudaError_t _e = (call); if(_e != cudaSuccess) return fail(ctx, #call, _e); } while(0)
#define CKP(call) do { cudaError_t _e = (call); if(_e != cudaSuccess){ fail(ctx, #call, _e); return nullptr; } } while(0)

extern "C" const char *bsb200_version(void){ return BSB_VERSION; }

extern "C" int bsb200_device_count(void){
	int n = 0;
	if(cudaGetDeviceCount(&n) != cudaSuccess) return 0;
	return n;
}

extern "C" bsb200_ctx *bsb200_create(int device, uint64_t trace_budget_bytes){
	int n = 0;
	if(cudaGetDeviceCount(&n) != cudaSuccess || n == 0 || device < 0 || device >= n){
		fprintf(stderr, "bsalign_b200: no usable CUDA device %d (count %d); this library has no CPU fallback\n", device, n);
		return nullptr;
	}
	bsb200_ctx *ctx = new bsb200_ctx();
	ctx->device = device;
	if(cudaSetDevice(device) != cudaSuccess){ delete ctx; return nullptr; }
	cudaDeviceProp prop;
	cudaGetDeviceProperties(&prop, device);
	ctx->num_sms = prop.multiProcessorCount;
	ctx->smem_optin = prop.sharedMemPerBlockOptin;
	ctx->total_mem = prop.totalGlobalMem;
	cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking);
	{ int lo = 0, hi = 0; cudaDeviceGetStreamPriorityRange(&lo, &hi); cudaStreamCreateWithPriority(&ctx->stream_bt, cudaStreamNonBlocking, hi); }
	for(auto &e : ctx->ev) cudaEventCreate(&e);
	for(auto &e : ctx->stage_ev) cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
	ctx->trace_budget = trace_budget_bytes;
	return ctx;
}

extern "C" void bsb200_destroy(bsb200_ctx *ctx){
	if(!ctx) return;
	if(ctx->helper){ bsb200_destroy(ctx->helper); ctx->helper = nullptr; }
	cudaSetDevice(ctx->device);
	cudaStreamSynchronize(ctx->stream);
	ctx->trace.release(); ctx->trace2.release(); ctx->counter.release();
	for(auto &d : ctx->dev_cache) d.release();
	for(auto &h : ctx->host_cache) h.release();
	for(auto &d : ctx->poa_cache) d.release();
	for(auto &d : ctx->kmer_cache) d.release();
	for(auto &d : ctx->remsa_cache) d.release();
	for(auto &h : ctx->poa_hcache) h.release();
	for(auto &e : ctx->ev) cudaEventDestroy(e);
	for(auto &e : ctx->stage_ev) if(e) cudaEventDestroy(e);
	for(auto &h : ctx->stage) h.release();
	cudaStreamDestroy(ctx->stream);
	cudaStreamDestroy(ctx->stream_bt);
	delete ctx;
}

// give the traceback arena and every cached buffer back to the device (the next batch allocates again)
extern "C" void bsb200_trim(bsb200_ctx *ctx){
	if(!ctx) return;
	if(ctx->helper) bsb200_trim(ctx->helper);
	cudaSetDevice(ctx->device);
	cudaStreamSynchronize(ctx->stream);
	ctx->trace.release(); ctx->trace2.release();
	for(auto &d : ctx->dev_cache) d.release();
	for(auto &h : ctx->host_cache) h.release();
	for(auto &d : ctx->poa_cache) d.release();
	for(auto &d : ctx->kmer_cache) d.release();
	for(auto &d : ctx->remsa_cache) d.release();
	for(auto &h : ctx->poa_hcache) h.release();
	for(auto &h : ctx->stage) h.release();
	ctx->auto_budget = 0;
}

// one shared context on device 0 for callers that have none (the compat headers; every translation unit gets the same one)
extern "C" bsb200_ctx *bsb200_default_context(void){
	static bsb200_ctx *ctx = nullptr;
	static std::atomic<int> lock{0};
	while(lock.exchange(1)) std::this_thread::yield();
	if(!ctx) ctx = bsb200_create(0, 0);
	lock.store(0);
	return ctx;
}

extern "C" const char *bsb200_last_error(bsb200_ctx *ctx){ return ctx ? ctx->err.c_str() : "null context"; }
extern "C" void bsb200_get_timing(bsb200_ctx *ctx, bsb200_timing_t *out){ if(ctx && out) *out = ctx->timing; }

extern "C" uint32_t bsb200_epi8_bandwidth(uint32_t qlen, uint32_t bandwidth){
	uint32_t bw = bandwidth ? bandwidth : qlen; // bsalign.h:3861-3862
	return (bw + 15) / 16 * 16;
}

extern "C" uint32_t bsb200_edit_bandwidth(uint32_t qlen, uint32_t tlen, int mode, uint32_t bandwidth){
	return edit_bandwidth(qlen, tlen, mode, bandwidth);
}

void bsb200_batch_free(bsb200_ctx *ctx, bsb200_batch *b);

// the O(n) loops of the host plan in slices on host threads (a million short pairs: the plan takes as long as their kernel)
static int plan_threads(uint64_t n){
	if(n < (1u << 16)) return 1;
	unsigned hc = std::thread::hardware_concurrency();
	if(const char *ev = getenv("BSB200_PLAN_THREADS")) hc = (unsigned)atoi(ev);
	return (int)std::max(1u, std::min(hc ? hc : 1u, 8u));   // (a thread costs ~25 us to start: 8 of them take a slice of 2-3 ms down to 0.5 ms)
}
template<class F> static void par_slices(uint64_t n, int nt, F fn){   // fn(slice, begin, end)
	if(nt <= 1){ fn(0, (uint64_t)0, n); return; }
	std::vector<std::thread> th;
	for(int w=1;w<nt;w++) th.emplace_back([=](){ fn(w, n * w / nt, n * (w + 1) / nt); });
	fn(0, (uint64_t)0, n / nt);
	for(auto &t : th) t.join();
}

// 2-bit packed sequences (the reference's BaseBank words, dna.h:63: base i = bits[i >> 5] >> (((~i) & 31) << 1) & 3) -> one base per byte
__global__ void __launch_bounds__(256) unpack_bits_kernel(const uint64_t *bits, uint8_t *out, uint64_t nwords, uint64_t nbases){
	const uint64_t w = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if(w >= nwords) return;
	const uint64_t v = bits[w];
	uint32_t o[8];
	#pragma unroll
	for(int k=0;k<8;k++){
		const uint32_t b4 = (uint32_t)(v >> (56 - 8 * k)) & 0xffu;   // bases 4k .. 4k+3, first base in the top two bits
		o[k] = ((b4 >> 6) & 3u) | (((b4 >> 4) & 3u) << 8) | (((b4 >> 2) & 3u) << 16) | ((b4 & 3u) << 24);
	}
	uint8_t *dst = out + w * 32;
	if(w * 32 + 32 <= nbases + 16){   // (the arena is allocated with 16 spare bytes and is 16-byte aligned)
		*(uint4*)dst = make_uint4(o[0], o[1], o[2], o[3]);
		if(w * 32 + 16 < nbases + 16) *(uint4*)(dst + 16) = make_uint4(o[4], o[5], o[6], o[7]);
	} else {
		for(uint64_t k=0;w*32+k<nbases;k++) dst[k] = (uint8_t)((o[k >> 2] >> (8 * (k & 3))) & 3u);
	}
}

// shared memory of one pair's group in the wavefront kernel: u, e and selector images + the scratch words of its stages
// ---- large copies between PAGEABLE caller memory and the device ----------------------------------------------------------------------
// cudaMemcpyAsync on pageable memory is staged by the driver on one thread (~8 GB/s here).  A reference caller holds plain malloc'ed
// arrays, so large copies are staged by us: host threads copy 32 MB pieces into / out of two pinned buffers while the DMA engine moves
// the other one.  Pinned or registered caller memory (and small copies) go straight to cudaMemcpyAsync.
static bool is_pageable(const void *p){
	cudaPointerAttributes at;
	if(cudaPointerGetAttributes(&at, p) != cudaSuccess){ cudaGetLastError(); return true; }
	return at.type == cudaMemoryTypeUnregistered;
}
// a call that cuts its work into chunks asks once per caller array and tells the copies of its chunks (-1: ask per copy)
static thread_local int tl_src_pageable = -1, tl_dst_pageable = -1;
constexpr size_t kStageChunk = 32ull << 20;
static void par_memcpy(void *dst, const void *src, size_t bytes){
	const int NT = bytes >= (8u << 20) ? 8 : 1;
	par_slices((bytes + 4095) / 4096, NT, [&](int, uint64_t lo, uint64_t hi){
		const size_t b0 = lo * 4096, b1 = std::min<size_t>(bytes, hi * 4096);
		if(b1 > b0) memcpy((uint8_t*)dst + b0, (const uint8_t*)src + b0, b1 - b0);
	});
}
static cudaError_t h2d_copy(bsb200_ctx *ctx, void *dst, const void *src, size_t bytes, cudaStream_t st){
	if(bytes < (4u << 20) || getenv("BSB200_NOSTAGE") || !(tl_src_pageable >= 0 ? tl_src_pageable != 0 : is_pageable(src))) return cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, st);
	cudaError_t e = cudaSuccess;
	for(int k=0;k<2;k++) if((e = ctx->stage[k].reserve(kStageChunk)) != cudaSuccess) return e;
	int k = 0;
	for(size_t off=0;off<bytes;off+=kStageChunk,k^=1){
		const size_t len = std::min(kStageChunk, bytes - off);
		if(off >= 2 * kStageChunk && (e = cudaEventSynchronize(ctx->stage_ev[k])) != cudaSuccess) return e;   // the DMA that last read this buffer
		par_memcpy(ctx->stage[k].p, (const uint8_t*)src + off, len);
		if((e = cudaMemcpyAsync((uint8_t*)dst + off, ctx->stage[k].p, len, cudaMemcpyHostToDevice, st)) != cudaSuccess) return e;
		cudaEventRecord(ctx->stage_ev[k], st);
	}
	// the staging buffers are reused by the next call: wait until the engine has read them
	for(int j=0;j<2;j++) if((e = cudaEventSynchronize(ctx->stage_ev[j])) != cudaSuccess) return e;
	return cudaSuccess;
}
// device -> pageable host; returns when the data is in dst (the stream is synchronised up to the copy)
static cudaError_t d2h_copy(bsb200_ctx *ctx, void *dst, const void *src, size_t bytes, cudaStream_t st){
	if(bytes < (4u << 20) || getenv("BSB200_NOSTAGE") || !(tl_dst_pageable >= 0 ? tl_dst_pageable != 0 : is_pageable(dst))) return cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, st);
	cudaError_t e = cudaSuccess;
	for(int k=0;k<2;k++) if((e = ctx->stage[k].reserve(kStageChunk)) != cudaSuccess) return e;
	const size_t nchunk = (bytes + kStageChunk - 1) / kStageChunk;
	for(size_t c=0;c<=nchunk;c++){
		if(c < nchunk){
			const size_t off = c * kStageChunk, len = std::min(kStageChunk, bytes - off);
			if((e = cudaMemcpyAsync(ctx->stage[c & 1].p, (const uint8_t*)src + off, len, cudaMemcpyDeviceToHost, st)) != cudaSuccess) return e;
			cudaEventRecord(ctx->stage_ev[c & 1], st);
		}
		if(c > 0){   // while chunk c is on the link, host threads move chunk c - 1 to its place
			const size_t off = (c - 1) * kStageChunk, len = std::min(kStageChunk, bytes - off);
			if((e = cudaEventSynchronize(ctx->stage_ev[(c - 1) & 1])) != cudaSuccess) return e;
			par_memcpy((uint8_t*)dst + off, ctx->stage[(c - 1) & 1].p, len);
		}
	}
	return cudaSuccess;
}

static size_t wave_group_smem(uint32_t max_bw, int split){
	const size_t img = epi8_image_bytes(max_bw / 16);
	const size_t scratch = std::max<size_t>(320, (size_t)(2 + 2 * kLanes * split) * 4);
	return (img * 3 + scratch + 127) / 128 * 128 + 32;
}

// Does the batch take the wavefront kernel, and with how many sub-blocks per lane?  Full bands (the band never moves), affine gaps
// in the ranges the kernel's unclamped adds need, scores inside +-63.  The split: the smallest one that gives an SM three warps per
// scheduler with the pairs that are in flight at once (bounded by the pairs, by shared memory and by the traceback store), among those
// that leave no sub-block of the narrowest band empty.
static int plan_wave(const bsb200_ctx *ctx, const bsb200_batch *b, uint32_t min_bw, uint64_t nact){
	if(b->kind != 0 || b->pw != 1 || getenv("BSB200_NOWAVE") || getenv("BSB200_NOFAST") || getenv("BSB200_NOFULL") || nact == 0) return 0;
	if(!(b->bandwidth == 0 || (b->bandwidth + 15) / 16 * 16 >= b->max_qlen)) return 0;
	for(int k=0;k<16;k++) if(b->mtx[k] > 63 || b->mtx[k] < -63) return 0;
	if(b->ge1 > -1 || (int)b->go1 + b->ge1 < -64 || (int)b->go1 + b->ge1 > -1) return 0;
	const uint64_t per_pair = ((uint64_t)epi8_row_bytes(b->max_bw / 16, 1) + kMetaInts * 4) * ((uint64_t)b->max_qlen + 64);
	const uint64_t by_mem = std::max<uint64_t>(1, (uint64_t)(ctx->total_mem * 0.85) / std::max<uint64_t>(per_pair, 1));
	int best = 0;
	for(int split : {1, 2, 4}){
		if(!epi8_wave_split_ok(min_bw / 16, (uint32_t)split)) break;
		if(split > 1 && !epi8_wave_use_anchors(b->max_bw / 16)) break;   // a traceback lookup must not cross sub-blocks: needs the sub-lane anchors
		const size_t gsm = wave_group_smem(b->max_bw, split);
		const size_t per_warp = gsm * (size_t)(4 / split);
		if(per_warp > ctx->smem_optin) continue;
		const uint64_t warps_by_smem = std::min<uint64_t>(64, (ctx->smem_optin + 1024) / (per_warp + 256)) * ctx->num_sms;
		const uint64_t in_flight = std::min<uint64_t>(std::min<uint64_t>(nact, by_mem), warps_by_smem * (4 / split));
		const double warps_per_sm = (double)in_flight * split / 4.0 / ctx->num_sms;
		best = split;
		if(warps_per_sm >= 12.0) break;
	}
	if(const char *ev = getenv("BSB200_WAVE_SPLIT")){   // tests / experiments: force a split (when the batch allows it)
		const int f = atoi(ev);
		if((f == 1 || f == 2 || f == 4) && best && (f == 1 || epi8_wave_use_anchors(b->max_bw / 16)) && epi8_wave_split_ok(min_bw / 16, (uint32_t)f) && wave_group_smem(b->max_bw, f) * (size_t)(4 / f) <= ctx->smem_optin) best = f;
	}
	return best;
}

// d_seqs_ext: the sequence arena is already in this device's memory (it arrived over NVLink: bsalign_b200/shard.py); the batch
// uses it in place and the caller keeps it alive until bsb200_batch_free.  Offsets and lengths are host arrays in both cases.
// bits64: the sequences come 2-bit packed (bsb200_batch_upload_bits): a quarter of the bytes cross PCIe, a kernel unpacks them.
// The sequences of a batch may come from TWO stretches of the caller's arena (all queries / all targets of a chunk of pairs): the
// second stretch starts at base `at1` of the batch's own arena; lengths in bases (bits: whole words, at1 a multiple of 32).
struct SeqSegs { uint64_t src1, len0, len1, at1; };

static bsb200_batch *upload_impl(bsb200_ctx *ctx, int kind, uint64_t n, const uint8_t *seqs, const uint8_t *d_seqs_ext, const uint64_t *bits64,
		const uint64_t *qoff, const uint32_t *qlen, const uint64_t *toff, const uint32_t *tlen,
		int mode, uint32_t bandwidth, const int8_t matrix[16], int8_t go1, int8_t ge1, int8_t go2, int8_t ge2, int want_cigar, const SeqSegs *segs = nullptr){
	if(!ctx) return nullptr;
	ctx->err.clear();
	if(n >= 0xFFFFFFF0ull || (n && ((!seqs && !d_seqs_ext && !bits64) || !qoff || !qlen || !toff || !tlen)) || (kind == 0 && !matrix) || kind < 0 || kind > 1){
		fail(ctx, "bsb200_batch_upload", cudaSuccess); return nullptr;
	}
	cudaSetDevice(ctx->device);
	// BSB200_HOSTPROF=1: wall-clock split of this call on stderr (development aid)
	const bool hostprof = getenv("BSB200_HOSTPROF") != nullptr;
	auto tp0 = std::chrono::steady_clock::now();
	auto lap = [&](const char *what){ if(hostprof){ auto t1 = std::chrono::steady_clock::now(); fprintf(stderr, "[upload] %-28s %8.3f ms\n", what, std::chrono::duration<double, std::milli>(t1 - tp0).count()); tp0 = t1; } };
	bsb200_batch *b = new bsb200_batch();
	{
		DevBuf *ds[] = {&b->d_seqs, &b->d_qoff, &b->d_toff, &b->d_qlen, &b->d_tlen, &b->d_order, &b->d_trace_off, &b->d_results, &b->d_status,
			&b->d_ncigar, &b->d_cig_raw, &b->d_cig_off, &b->d_cig_dense, &b->d_dense_off, &b->d_dense_total, &b->d_block_rows, &b->d_prefix, &b->d_bits};
		for(int k=0;k<18;k++){ *ds[k] = ctx->dev_cache[k]; ctx->dev_cache[k] = DevBuf(); }
		HostBuf *hs[] = {&b->h_results, &b->h_status, &b->h_ncigar, &b->h_dense_off, &b->h_dense, &b->h_total};
		for(int k=0;k<6;k++){ *hs[k] = ctx->host_cache[k]; ctx->host_cache[k] = HostBuf(); }
	}
	b->kind = kind; b->n = n; b->mode = mode & 3; b->bandwidth = bandwidth; b->want_cigar = want_cigar;
	if(kind == 0){
		memcpy(b->mtx, matrix, 16); b->go1 = go1; b->ge1 = ge1; b->go2 = go2; b->ge2 = ge2;
		b->pw = epi8_piecewise(go1, ge1, go2, ge2, 16);
	}
	// ---- the caller's arrays start crossing PCIe before the host plans: the copies do not depend on the plan ----------------
	size_t seq_end = 0;
	uint64_t cig_cap_words = 0;   // per-pair cigar capacity (qlen + tlen + 2 words), summed: sizes the side buffers
	const int NT = plan_threads(n);
	{
		std::vector<uint64_t> pe(NT, 0), pc(NT, 0);
		par_slices(n, NT, [&](int w, uint64_t lo, uint64_t hi){
			uint64_t se = 0, cc = 0;
			for(uint64_t i=lo;i<hi;i++){
				se = std::max<uint64_t>(se, std::max<uint64_t>(qoff[i] + qlen[i], toff[i] + tlen[i]));
				if(qlen[i] && tlen[i]) cc += (uint64_t)qlen[i] + tlen[i] + 2;
			}
			pe[w] = se; pc[w] = cc;
		});
		for(int w=0;w<NT;w++){ seq_end = std::max<size_t>(seq_end, pe[w]); cig_cap_words += pc[w]; }
	}
	cudaStream_t st = ctx->stream;
	cudaError_t e = cudaSuccess;
	auto R = [&](cudaError_t x){ if(e == cudaSuccess) e = x; };
	// every early return below first waits for the copies in flight (they read the caller's memory)
	auto bail = [&]() -> bsb200_batch* { cudaStreamSynchronize(st); bsb200_batch_free(ctx, b); return nullptr; };
	if(!d_seqs_ext) R(b->d_seqs.reserve(seq_end + 16));
	b->d_seqs_ext = d_seqs_ext;
	R(b->d_qoff.reserve(n * 8 + 8)); R(b->d_toff.reserve(n * 8 + 8));
	R(b->d_qlen.reserve(n * 4 + 4)); R(b->d_tlen.reserve(n * 4 + 4));
	if(e != cudaSuccess){ fail(ctx, "device allocation", e); return bail(); }
	// (a first attempt ran the plan of a million edit pairs three times slower next to the DMA stream - config 4 end to end
	// 53 -> 93 ms; once the plan no longer touched 50 MB of fresh per-pair arrays the overlap paid off: 35.6 -> 28.1 ms)
	const bool early = true;
	auto copy_inputs = [&](){
		cudaEventRecord(ctx->ev[0], st);
		if(n){
			if(bits64){
				const uint64_t nw = (seq_end + 31) / 32;
				R(b->d_bits.reserve(nw * 8 + 8));
				if(e == cudaSuccess){
					if(segs){
						R(cudaMemcpyAsync(b->d_bits.p, bits64, (segs->len0 + 31) / 32 * 8, cudaMemcpyHostToDevice, st));
						R(cudaMemcpyAsync(b->d_bits.as<uint64_t>() + segs->at1 / 32, bits64 + segs->src1 / 32, (segs->len1 + 31) / 32 * 8, cudaMemcpyHostToDevice, st));
					} else R(h2d_copy(ctx, b->d_bits.p, bits64, nw * 8, st));
					unpack_bits_kernel<<<(unsigned)((nw + 255) / 256), 256, 0, st>>>(b->d_bits.as<uint64_t>(), b->d_seqs.as<uint8_t>(), nw, seq_end);
				}
			} else if(!d_seqs_ext){
				if(segs){
					R(h2d_copy(ctx, b->d_seqs.p, seqs, segs->len0, st));
					R(h2d_copy(ctx, b->d_seqs.as<uint8_t>() + segs->at1, seqs + segs->src1, segs->len1, st));
				} else R(h2d_copy(ctx, b->d_seqs.p, seqs, seq_end, st));
			}
			R(h2d_copy(ctx, b->d_qoff.p, qoff, n * 8, st));
			R(h2d_copy(ctx, b->d_toff.p, toff, n * 8, st));
			R(h2d_copy(ctx, b->d_qlen.p, qlen, n * 4, st));
			R(h2d_copy(ctx, b->d_tlen.p, tlen, n * 4, st));
		}
	};
	if(early) copy_inputs();
	lap("buffers + seq_end + early H2D");
	// ---- plan: per-pair band, trace footprint, heaviest-first order, waves ---------------------------
	std::vector<uint64_t> work(kind == 0 ? n : 0), tbytes(kind == 0 ? n : 0);   // epi8 only (a million short edit pairs: every O(n) array counts)
	b->max_bw = kind == 0 ? 16 : 64;
	uint32_t wave_slack = 0;
	if(kind == 0){
		// the wavefront kernel's skewed trace needs extra slots per pair, how many depends on the split: decide it before sizing
		uint32_t min_bw = 0xffffffffu; uint64_t nact0 = 0;
		for(uint64_t i=0;i<n;i++){
			if(qlen[i] == 0 || tlen[i] == 0) continue;
			const uint32_t bw = bsb200_epi8_bandwidth(qlen[i], bandwidth);
			min_bw = std::min(min_bw, bw); b->max_bw = std::max(b->max_bw, bw); b->max_qlen = std::max(b->max_qlen, qlen[i]);
			nact0++;
		}
		b->wave_split = plan_wave(ctx, b, min_bw, nact0);
		wave_slack = b->wave_split ? epi8_wave_slack((uint32_t)b->wave_split) : 0;
	}
	b->empty.assign(n, 0);
	{
		struct Part { uint64_t cells = 0, trace = 0, nact = 0, nlong = 0; uint32_t max_q64 = 0, max_bw = 0, max_qlen = 0; };
		std::vector<Part> part(NT);
		par_slices(n, NT, [&](int w, uint64_t lo, uint64_t hi){
			Part p;
			for(uint64_t i=lo;i<hi;i++){
				uint32_t bw;
				if(qlen[i] == 0 || tlen[i] == 0){ b->empty[i] = 1; continue; } // bsalign.h:1051-1054 (work/tbytes stay 0)
				if(kind == 0){
					bw = bsb200_epi8_bandwidth(qlen[i], bandwidth);
					// upper bound: with sub-lane anchors, and with the extra slots of the wavefront kernel's skewed layout
					tbytes[i] = ((uint64_t)epi8_row_bytes(bw / 16, b->pw) + kMetaInts * 4) * ((uint64_t)tlen[i] + 1 + wave_slack);
					tbytes[i] = (tbytes[i] + 127) / 128 * 128;   // every pair's block starts on a 128-byte line: the chunk stores are whole sectors
					p.cells += (uint64_t)std::min<uint32_t>(bw, (qlen[i] + 15) / 16 * 16) * tlen[i];
					p.trace += ((uint64_t)bw * (b->pw + 1) + 84) * tlen[i];
					work[i] = (uint64_t)bw * tlen[i];
				} else {
					bw = edit_bandwidth(qlen[i], tlen[i], mode, bandwidth);
					{   // beyond edit_kernel: a band over 16384 cells, or query bit-planes that outgrow shared memory even for one warp per CTA
						const uint32_t q64 = (qlen[i] + 63) / 64 * 64;
						if(bw > 16384 || (uint64_t)(q64 / 64 + 2) * 16 * 32 > ctx->smem_optin){ b->empty[i] = bw == q64 ? 2 : 3; p.nlong++; continue; }
					}
					p.cells += (uint64_t)bw * tlen[i];
					p.trace += ((uint64_t)bw / 4 + 4) * tlen[i];
					p.max_q64 = std::max<uint32_t>(p.max_q64, (qlen[i] + 63) / 64 * 64);
				}
				p.max_bw = std::max(p.max_bw, bw);
				p.max_qlen = std::max(p.max_qlen, qlen[i]);
				p.nact++;
			}
			part[w] = p;
		});
		uint64_t nact_ = 0, nlong_ = 0;
		for(auto &p : part) nlong_ += p.nlong;
		if(nlong_) for(uint64_t i=0;i<n;i++) if(b->empty[i] == 2){   // scratch of edit_long_kernel: query planes + 2 planes x (tlen + 1) rows
			b->long_pairs.push_back((uint32_t)i); b->long_off.push_back(b->long_bytes);
			const uint64_t W = ((uint64_t)qlen[i] + 63) / 64;
			b->long_bytes += (16 * W * ((uint64_t)tlen[i] + 2) + 127) / 128 * 128;
			b->cells += W * 64 * tlen[i];
		}
		for(auto &p : part){
			b->cells += p.cells; b->trace_bytes += p.trace; nact_ += p.nact;
			b->max_q64 = std::max(b->max_q64, p.max_q64); b->max_bw = std::max(b->max_bw, p.max_bw); b->max_qlen = std::max(b->max_qlen, p.max_qlen);
		}
		// the active pairs in input order (the sort below is stable on it): every slice writes behind the slices before it
		b->order.resize(nact_);
		std::vector<uint64_t> start(NT + 1, 0);
		for(int w=0;w<NT;w++) start[w + 1] = start[w] + part[w].nact;
		par_slices(n, NT, [&](int w, uint64_t lo, uint64_t hi){
			uint64_t k = start[w];
			for(uint64_t i=lo;i<hi;i++) if(!b->empty[i]) b->order[k++] = (uint32_t)i;
		});
	}
	b->seq_bytes = seq_end;
	const uint32_t nact = (uint32_t)b->order.size();
	lap("per-pair loop");
	{
		// heaviest first (epi8: band cells; edit: target length, so the 32 pairs of a warp finish together); counting
		// sort when the key range is small, else one sort of packed (key, index) words
		// the order only balances the schedule: epi8 keys are quantised to 16 bits so that the counting sort's table stays small;
		// the key is recomputed where it is used instead of being stored (one O(n) array less)
		uint64_t kmax = 0;
		{
			std::vector<uint64_t> pk(NT, 0);
			par_slices(nact, NT, [&](int w, uint64_t lo, uint64_t hi){ uint64_t m = 0; for(uint64_t k=lo;k<hi;k++){ uint32_t i = b->order[k]; m = std::max<uint64_t>(m, kind == 0 ? work[i] : (uint64_t)tlen[i]); } pk[w] = m; });
			for(int w=0;w<NT;w++) kmax = std::max(kmax, pk[w]);
		}
		int sh = 0;
		if(kind == 0) while((kmax >> sh) >= (1u << 16)) sh++;
		kmax >>= sh;
		auto key_of = [&](uint32_t i) -> uint64_t { return kind == 0 ? (work[i] >> sh) : (uint64_t)tlen[i]; };
		if(kmax < (1u << 22)){
			// stable counting sort, descending key; slices count their own keys, then every (key, slice) gets its place
			const int ST = (nact >= (1u << 16) && kmax < (1u << 17)) ? NT : 1;
			std::vector<std::vector<uint32_t>> cnt(ST, std::vector<uint32_t>(kmax + 1, 0));
			par_slices(nact, ST, [&](int w, uint64_t lo, uint64_t hi){ for(uint64_t k=lo;k<hi;k++) cnt[w][kmax - key_of(b->order[k])]++; });
			uint32_t run = 0;
			for(uint64_t v=0;v<=kmax;v++) for(int w=0;w<ST;w++){ const uint32_t c = cnt[w][v]; cnt[w][v] = run; run += c; }
			std::vector<uint32_t> sorted(nact);
			par_slices(nact, ST, [&](int w, uint64_t lo, uint64_t hi){ for(uint64_t k=lo;k<hi;k++){ const uint32_t i = b->order[k]; sorted[cnt[w][kmax - key_of(i)]++] = i; } });
			b->order.swap(sorted);
		} else {
			std::vector<std::pair<uint64_t, uint32_t>> kv(nact);
			for(uint32_t k=0;k<nact;k++) kv[k] = std::make_pair(~key_of(b->order[k]), b->order[k]);
			std::sort(kv.begin(), kv.end());
			for(uint32_t k=0;k<nact;k++) b->order[k] = kv[k].second;
		}
	}
	lap("sort");
	// Budget of the traceback arena per wave.  Automatic (trace_budget 0): 90 % of what is free after this batch's other device
	// buffers (sequences, tables, results, two cigar arenas; counted in full although cached ones are re-used) and the same again
	// for a second batch on this context: batches whose pairs are bound by a dependency chain (10 kb bands: one warp's issue rate
	// per pair) gain throughput only through the pairs a wave seats.  cudaMemGetInfo takes milliseconds once a 100+ GB arena
	// exists, so an existing arena is planned against first and the device is only asked when that splits the batch and the last
	// answer promised noticeably more.
	auto query_budget = [&]() -> uint64_t {
		size_t fr = 0, tot = 0;
		cudaMemGetInfo(&fr, &tot);
		const uint64_t side = seq_end + n * 100 + (want_cigar ? cig_cap_words * 8 + n * 8 : 0) + (64ull << 20);
		const uint64_t avail = (uint64_t)fr + ctx->trace.cap + ctx->trace2.cap;
		ctx->auto_budget = avail > 2 * side ? (uint64_t)((avail - 2 * side) * 0.90) : avail / 2;
		return ctx->auto_budget;
	};
	uint64_t budget = ctx->trace_budget;
	bool queried = budget != 0;
	if(budget == 0){
		if(ctx->trace.cap > 512) budget = ctx->trace.cap - 512;
		else { budget = query_budget(); queried = true; }
	}
	const uint64_t ntoff = kind == 0 ? n : ((uint64_t)nact + 31) / 32;   // epi8: per pair; edit: per block of 32 pairs
	b->cig_off.assign(n + 1, 0);
	if(want_cigar){
		std::vector<uint64_t> base(NT + 1, 0);
		par_slices(n, NT, [&](int w, uint64_t lo, uint64_t hi){ uint64_t c = 0; for(uint64_t i=lo;i<hi;i++) c += (qlen[i] && tlen[i]) ? (uint64_t)qlen[i] + tlen[i] + 2 : 0; base[w + 1] = c; });
		for(int w=0;w<NT;w++) base[w + 1] += base[w];
		par_slices(n, NT, [&](int w, uint64_t lo, uint64_t hi){ uint64_t c = base[w]; for(uint64_t i=lo;i<hi;i++){ c += (qlen[i] && tlen[i]) ? (uint64_t)qlen[i] + tlen[i] + 2 : 0; b->cig_off[i + 1] = c; } });
		b->cig_words = b->cig_off[n];
	}
	lap("cig_off");
	// ---- device buffers: everything but the traceback arena first, the arena takes what is left ---------------------
	R(b->d_order.reserve(n * 4 + 4));
	R(b->d_trace_off.reserve(n * 8 + 8)); R(b->d_results.reserve(n * 40 + 40)); R(b->d_status.reserve(n * 4 + 4));
	R(b->d_ncigar.reserve(n * 4 + 4)); R(b->d_dense_off.reserve(n * 8 + 8)); R(b->d_dense_total.reserve(16));
	if(want_cigar){ R(b->d_cig_raw.reserve(b->cig_words * 4 + 16)); R(b->d_cig_off.reserve((n + 1) * 8)); R(b->d_cig_dense.reserve(b->cig_words * 4 + 16)); }
	R(ctx->counter.reserve(256));
	if(e != cudaSuccess){ fail(ctx, "device allocation", e); return bail(); }
	lap("side buffers");
	std::vector<uint32_t> block_rows, blk_max;
	// ---- waves against the budget; when the arena of an automatic budget cannot be had after all (memory taken by someone else
	// since the query), plan again with three quarters of it ---------------------------------------------------------------
	for(int attempt=0;;attempt++){
		b->waves.clear();
		b->trace_off.assign(ntoff + 1, 0);
		if(kind == 0){
			// as few waves as the trace budget allows: a wave must keep every SM's groups busy, and splitting further to
			// overlap the traceback with the next forward sweep cost more (idle groups, co-scheduling) than it hid
			Wave w = {0, 0, 0};
			for(uint32_t k=0;k<nact;k++){
				uint32_t i = b->order[k];
				if(tbytes[i] > budget){
					if(!queried){ queried = true; const uint64_t bq = query_budget(); if(bq > budget){ budget = bq; w.end = w.beg = 0; k = (uint32_t)-1; b->waves.clear(); w.trace_bytes = 0; continue; } }
					ctx->err = "a single pair needs more traceback memory than the budget"; return bail();
				}
				if(w.trace_bytes + tbytes[i] > budget && w.end > w.beg){
					b->waves.push_back(w);
					w.beg = w.end; w.trace_bytes = 0;
				}
				b->trace_off[i] = w.trace_bytes;
				w.trace_bytes += tbytes[i];
				w.end = k + 1;
			}
			if(w.end > w.beg) b->waves.push_back(w);
		} else {
			// 32 consecutive pairs (one warp) share an interleaved trace block sized by the longest target in it
			const uint32_t WB = b->max_bw / 64;
			const uint32_t nblk = (nact + 31) / 32;
			block_rows.assign(nblk + 1, 0);
			if(attempt == 0 || blk_max.size() != nblk){
				blk_max.assign(nblk, 0);
				par_slices(nblk, NT, [&](int, uint64_t lo_, uint64_t hi_){
					for(uint64_t bk=lo_;bk<hi_;bk++){
						uint32_t lo = (uint32_t)bk * 32, hi = std::min<uint32_t>(lo + 32, nact), mt = 0;
						for(uint32_t k=lo;k<hi;k++) mt = std::max(mt, tlen[b->order[k]]);
						blk_max[bk] = mt;
					}
				});
			}
			Wave w = {0, 0, 0};
			for(uint32_t bk=0;bk<nblk;bk++){
				uint32_t lo = bk * 32, hi = std::min<uint32_t>(lo + 32, nact);
				(void)lo;
				uint64_t R_ = (uint64_t)blk_max[bk] + 1;
				uint64_t bytes = R_ * WB * 2 * 32 * 8 + R_ * 32 * 4;
				if(bytes > budget){
					if(!queried){ queried = true; const uint64_t bq = query_budget(); if(bq > budget){ budget = bq; w.end = w.beg = 0; bk = (uint32_t)-1; b->waves.clear(); w.trace_bytes = 0; continue; } }
					ctx->err = "a single block of pairs needs more traceback memory than the budget"; return bail();
				}
				if(w.trace_bytes + bytes > budget && w.end > w.beg){
					b->waves.push_back(w);
					w.beg = w.end; w.trace_bytes = 0;
				}
				block_rows[bk] = (uint32_t)R_;
				b->trace_off[bk] = w.trace_bytes;
				w.trace_bytes += bytes;
				w.end = hi;
			}
			if(w.end > w.beg) b->waves.push_back(w);
		}
		if(!queried && b->waves.size() > 1 && (ctx->auto_budget == 0 || ctx->auto_budget > budget + budget / 4)){
			queried = true;
			const uint64_t bq = query_budget();
			if(bq > budget + budget / 4){ budget = bq; continue; }   // a noticeably larger arena can be had: plan again
		}
		uint64_t max_wave = 0;
		for(auto &w : b->waves) max_wave = std::max(max_wave, w.trace_bytes);
		b->max_wave_bytes = max_wave;
		const cudaError_t ea = ctx->trace.reserve(max_wave + 256, true);
		if(ea == cudaSuccess) break;
		cudaGetLastError();   // clear the sticky allocation error
		if(ctx->trace_budget != 0 || attempt >= 3){ fail(ctx, "device allocation (traceback arena)", ea); return bail(); }
		budget = budget / 4 * 3;
	}
	lap("waves + arena");
	if(kind == 1) R(b->d_block_rows.reserve(block_rows.size() * 4 + 16));
	if(e != cudaSuccess){ fail(ctx, "device allocation", e); return bail(); }
	if(!early) copy_inputs();
	if(n){
		if(nact) R(cudaMemcpyAsync(b->d_order.p, b->order.data(), (size_t)nact * 4, cudaMemcpyHostToDevice, st));
		if(kind == 1 && !block_rows.empty()) R(cudaMemcpyAsync(b->d_block_rows.p, block_rows.data(), block_rows.size() * 4, cudaMemcpyHostToDevice, st));
		if(ntoff) R(cudaMemcpyAsync(b->d_trace_off.p, b->trace_off.data(), ntoff * 8, cudaMemcpyHostToDevice, st));
		if(want_cigar) R(cudaMemcpyAsync(b->d_cig_off.p, b->cig_off.data(), (n + 1) * 8, cudaMemcpyHostToDevice, st));
	}
	cudaEventRecord(ctx->ev[1], st);
	R(cudaStreamSynchronize(st)); // order/trace_off are host vectors owned by b, but seqs belong to the caller
	if(e != cudaSuccess){ fail(ctx, "host to device copy", e); return bail(); }
	lap("H2D + sync");
	float ms = 0; cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[1]);
	ctx->timing = bsb200_timing_t();
	ctx->timing.h2d_ms = ms;
	{
		const uint64_t sb = segs ? segs->len0 + segs->len1 : seq_end;
		ctx->timing.h2d_bytes = (d_seqs_ext ? 0 : (bits64 ? (sb + 31) / 32 * 8 : sb)) + n * (8 + 8 + 4 + 4 + 4 + 8) + (want_cigar ? (n + 1) * 8 : 0);
	}
	ctx->timing.cells = b->cells;
	ctx->timing.trace_bytes = b->trace_bytes;
	ctx->timing.waves = (uint32_t)b->waves.size();
	return b;
}

extern "C" bsb200_batch *bsb200_batch_upload(bsb200_ctx *ctx, int kind, uint64_t n, const uint8_t *seqs,
		const uint64_t *qoff, const uint32_t *qlen, const uint64_t *toff, const uint32_t *tlen,
		int mode, uint32_t bandwidth, const int8_t matrix[16], int8_t go1, int8_t ge1, int8_t go2, int8_t ge2, int want_cigar){
	return upload_impl(ctx, kind, n, seqs, nullptr, nullptr, qoff, qlen, toff, tlen, mode, bandwidth, matrix, go1, ge1, go2, ge2, want_cigar);
}

extern "C" bsb200_batch *bsb200_batch_upload_dev(bsb200_ctx *ctx, int kind, uint64_t n, const uint8_t *d_seqs,
		const uint64_t *qoff, const uint32_t *qlen, const uint64_t *toff, const uint32_t *tlen,
		int mode, uint32_t bandwidth, const int8_t matrix[16], int8_t go1, int8_t ge1, int8_t go2, int8_t ge2, int want_cigar){
	return upload_impl(ctx, kind, n, nullptr, d_seqs, nullptr, qoff, qlen, toff, tlen, mode, bandwidth, matrix, go1, ge1, go2, ge2, want_cigar);
}

extern "C" bsb200_batch *bsb200_batch_upload_bits(bsb200_ctx *ctx, int kind, uint64_t n, const uint64_t *bits,
		const uint64_t *qoff, const uint32_t *qlen, const uint64_t *toff, const uint32_t *tlen,
		int mode, uint32_t bandwidth, const int8_t matrix[16], int8_t go1, int8_t ge1, int8_t go2, int8_t ge2, int want_cigar){
	return upload_impl(ctx, kind, n, nullptr, nullptr, bits, qoff, qlen, toff, tlen, mode, bandwidth, matrix, go1, ge1, go2, ge2, want_cigar);
}

// row store of the two-pass kernel: cp.async.bulk copies out of the shared-memory images (1) or the LDS.128 -> STG.128 loop (0).
// Measured (profiles/ab_bulk_store_r2.txt): full bands (config 2 on this kernel, 1 KB images) 18.80 -> 17.33 ms with the bulk copies;
// moving bands (config 3, 512-byte images, the copy has only the steering code to hide behind) 80.93 -> 81.14 ms.  Hence on for
// full-band batches only; BSB200_BULK_STORE overrides (the A/B).
static int bulk_store_default(bool full){
	if(const char *ev = getenv("BSB200_BULK_STORE")) return atoi(ev) != 0;
	return full ? 1 : 0;
}

template<int PW, bool FAST, bool ANCH>
static int launch_epi8_forward(bsb200_ctx *ctx, Epi8Args a, uint32_t npairs, bool full){
	// CTA shape: as many groups per SM as the shared memory allows.  Normally 128 threads = 4 warps x 4 groups; a wide band
	// (tens of KB per group) does better with one-warp CTAs that run fewer than 4 groups each.
	const size_t sm_bytes = ctx->smem_optin + 1024;   // per-SM shared memory ~ opt-in maximum per block + its reserved KB
	int best_threads = 0; uint32_t best_gpw = 0, best_groups = 0;
	double best_cost = 1e30;
	const int shapes[6][2] = {{128, 4}, {64, 4}, {32, 4}, {32, 3}, {32, 2}, {32, 1}};
	for(auto &sh : shapes){
		size_t per_cta = (size_t)(sh[0] / 32) * sh[1] * a.group_smem;
		if(per_cta > ctx->smem_optin) continue;
		uint32_t ctas = (uint32_t)std::min<size_t>(sm_bytes / (per_cta + 1024), 2048 / sh[0]);
		ctas = std::min<uint32_t>(ctas, 32);
		uint32_t groups = ctas * (sh[0] / 32) * sh[1];
		if(!groups) continue;
		// rounds needed to seat every pair x relative cost of a round (a warp instruction serves gpw groups; measured
		// on 10 kb bands: a round at 1 group per warp takes ~1.75x a round at 4)
		uint64_t slots = (uint64_t)groups * ctx->num_sms;
		double cost = (double)((npairs + slots - 1) / slots) * (1.0 + 0.25 * (4 - sh[1]));
		if(cost < best_cost - 1e-9 || (cost < best_cost + 1e-9 && groups > best_groups)){ best_cost = cost; best_groups = groups; best_threads = sh[0]; best_gpw = sh[1]; }
	}
	if(!best_groups){ ctx->err = "band too wide for the shared-memory row buffers"; return -1; }
	if constexpr (ANCH){
		// test hook: one-warp CTAs with 1..3 groups per warp (the NARROW instantiations) on any batch that carries anchors
		if(const char *ev = getenv("BSB200_GPW")){
			int g = atoi(ev);
			if(g >= 1 && g <= 3 && (size_t)g * a.group_smem <= ctx->smem_optin){ best_threads = 32; best_gpw = (uint32_t)g; best_groups = (uint32_t)g; }
		}
	}
	a.gpw = best_gpw;
	const int threads = best_threads;
	const size_t smem = (size_t)(threads / 32) * best_gpw * a.group_smem;
	auto go = [&](auto kernel) -> int {
		CK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
		int per_sm = 1;
		CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, smem));
		if(per_sm < 1) per_sm = 1;
		uint32_t groups = (threads / 32) * best_gpw;
		uint32_t grid = (npairs + groups - 1) / groups;
		// a grid that seats every pair at once is rounded up to whole CTAs-per-SM: groups race for pairs (atomic counter), so each SM ends
		// up with about the same number of pairs instead of some SMs carrying one CTA more than the others (config 3: 625 CTAs on 148 SMs)
		if(grid > (uint32_t)ctx->num_sms) grid = (grid + ctx->num_sms - 1) / ctx->num_sms * ctx->num_sms;
		grid = std::min<uint32_t>(grid, (uint32_t)(ctx->num_sms * per_sm));
		if(grid == 0) grid = 1;
		kernel<<<grid, threads, smem, ctx->stream>>>(a);
		CK(cudaGetLastError());
		return 0;
	};
	// batches that leave an SM with at most two warps per scheduler are bound by the latency of the row loop's dependency chain, not
	// by ALU throughput: they take the LAT instantiation (short F chain, loads of the next chunk in flight; affine gaps only)
	bool lat = false;
	if constexpr (FAST && PW == 1){
		const uint64_t seated = std::min<uint64_t>(npairs, (uint64_t)best_groups * ctx->num_sms);
		const uint64_t warps_per_sm = (seated + (uint64_t)best_gpw * ctx->num_sms - 1) / ((uint64_t)best_gpw * ctx->num_sms);
		lat = warps_per_sm <= 8;
		if(const char *ev = getenv("BSB200_LAT")) lat = atoi(ev) != 0;   // experiments
	}
	if(best_gpw < 4){
		if constexpr (ANCH){
			if constexpr (FAST && PW == 1){
				if(full) return lat ? go(epi8_forward_kernel<PW, FAST, true, true, true, true>) : go(epi8_forward_kernel<PW, FAST, true, true, false, true>);
				if(lat) return go(epi8_forward_kernel<PW, FAST, true, true, true>);
			}
			return go(epi8_forward_kernel<PW, FAST, true, true>);
		} else { ctx->err = "internal: narrow warps without anchors"; return -1; }
	}
	if constexpr (FAST && PW == 1){
		// full-band batches (bandwidth 0) never move their band: instantiation without the shift / steering code
		if(full) return lat ? go(epi8_forward_kernel<PW, FAST, ANCH, false, true, true>) : go(epi8_forward_kernel<PW, FAST, ANCH, false, false, true>);
		if(lat) return go(epi8_forward_kernel<PW, FAST, ANCH, false, true>);
	}
	return go(epi8_forward_kernel<PW, FAST, ANCH, false>);
}

// the wavefront kernel: CTAs of whole warps (4 / split pairs each), as many pairs per SM as the shared memory seats
template<bool ANCH, int SPLIT>
static int launch_epi8_wave_t(bsb200_ctx *ctx, Epi8Args a, uint32_t npairs){
	const size_t per_warp = (size_t)a.group_smem * (4 / SPLIT);
	int threads = 0; uint32_t best_pairs = 0;
	for(int th : {128, 64, 32}){
		const size_t per_cta = per_warp * (th / 32);
		if(per_cta > ctx->smem_optin) continue;
		const uint32_t ctas = (uint32_t)std::min<size_t>((ctx->smem_optin + 1024) / (per_cta + 1024), 2048 / th);
		const uint32_t pairs_sm = std::min<uint32_t>(ctas, 32) * (th / 32) * (4 / SPLIT);
		if(pairs_sm > best_pairs){ best_pairs = pairs_sm; threads = th; }
	}
	if(!threads){ ctx->err = "band too wide for the shared-memory row buffers"; return -1; }
	const size_t smem = per_warp * (threads / 32);
	auto kernel = epi8_wave_kernel<ANCH, SPLIT>;
	CK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
	int per_sm = 1;
	CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, smem));
	if(per_sm < 1) per_sm = 1;
	const uint32_t groups = (threads / 32) * (4 / SPLIT);
	uint32_t grid = (npairs + groups - 1) / groups;
	if(grid > (uint32_t)ctx->num_sms) grid = (grid + ctx->num_sms - 1) / ctx->num_sms * ctx->num_sms;
	grid = std::min<uint32_t>(grid, (uint32_t)(ctx->num_sms * per_sm));
	if(grid == 0) grid = 1;
	kernel<<<grid, threads, smem, ctx->stream>>>(a);
	CK(cudaGetLastError());
	return 0;
}

static int launch_epi8_wave(bsb200_ctx *ctx, const Epi8Args &a, uint32_t npairs, bool anch, int split){
	if(anch){
		if(split == 4) return launch_epi8_wave_t<true, 4>(ctx, a, npairs);
		if(split == 2) return launch_epi8_wave_t<true, 2>(ctx, a, npairs);
		return launch_epi8_wave_t<true, 1>(ctx, a, npairs);
	}
	if(split != 1){ ctx->err = "internal: split wavefront kernel without anchors"; return -1; }
	return launch_epi8_wave_t<false, 1>(ctx, a, npairs);
}

template<bool FAST, bool ANCH>
static int launch_epi8_forward_pw(bsb200_ctx *ctx, const Epi8Args &a, uint32_t npairs, int pw, bool full){
	if(pw == 2) return launch_epi8_forward<2, FAST, ANCH>(ctx, a, npairs, full);
	if(pw == 1) return launch_epi8_forward<1, FAST, ANCH>(ctx, a, npairs, full);
	return launch_epi8_forward<0, FAST, ANCH>(ctx, a, npairs, full);
}

template<int WR>
static int launch_edit_t(bsb200_ctx *ctx, const EditArgs &a){
	int threads = kEditThreads;
	size_t per_thread = (size_t)a.nQW * 16;
	while(threads > 32 && per_thread * threads > ctx->smem_optin) threads >>= 1;
	if(per_thread * threads > ctx->smem_optin){ ctx->err = "query too long for the shared-memory bit-planes of the edit kernel"; return -1; }
	size_t smem = per_thread * threads;
	CK(cudaFuncSetAttribute(edit_kernel<WR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
	uint32_t grid = (a.npairs + threads - 1) / threads;
	edit_kernel<WR><<<grid, threads, smem, ctx->stream>>>(a);
	CK(cudaGetLastError());
	return 0;
}

static int launch_edit(bsb200_ctx *ctx, const EditArgs &a){
	if(a.WB <= 1) return launch_edit_t<1>(ctx, a);
	if(a.WB <= 2) return launch_edit_t<2>(ctx, a);
	if(a.WB <= 4) return launch_edit_t<4>(ctx, a);
	if(a.WB <= 8) return launch_edit_t<8>(ctx, a);
	if(a.WB <= 16) return launch_edit_t<16>(ctx, a);
	if(a.WB <= 256) return launch_edit_t<256>(ctx, a);
	ctx->err = "edit band wider than 16384 cells is not supported";
	return -1;
}

extern "C" int bsb200_batch_run(bsb200_ctx *ctx, bsb200_batch *b){
	if(!ctx || !b) return -1;
	ctx->err.clear();
	cudaSetDevice(ctx->device);
	cudaStream_t st = ctx->stream;
	ctx->timing.forward_launches = ctx->timing.traceback_launches = ctx->timing.other_launches = 0;
	float fwd_ms = 0, bt_ms = 0;
	// four events per wave, destroyed on every way out (an error return first waits for the kernels already in flight)
	struct EventSet {
		std::vector<cudaEvent_t> v; cudaStream_t st;
		~EventSet(){ if(!v.empty()) cudaStreamSynchronize(st); for(auto &e : v) if(e) cudaEventDestroy(e); }
		cudaEvent_t &operator[](size_t i){ return v[i]; }
	} evs;
	evs.st = st;
	if(b->n == 0 || (b->waves.empty() && b->long_pairs.empty())){ if(b->n){ cudaMemsetAsync(b->d_results.p, 0, b->n * 40, st); cudaMemsetAsync(b->d_status.p, 0, b->n * 4, st); cudaMemsetAsync(b->d_ncigar.p, 0, b->n * 4, st); cudaMemsetAsync(b->d_dense_off.p, 0, b->n * 8, st); cudaMemsetAsync(b->d_dense_total.p, 0, 16, st); cudaStreamSynchronize(st);} b->ran = true; return 0; }
	CK(cudaEventRecord(ctx->ev[4], st));
	CK(cudaMemsetAsync(b->d_results.p, 0, b->n * 40, st));
	CK(cudaMemsetAsync(b->d_status.p, 0, b->n * 4, st));
	CK(cudaMemsetAsync(b->d_ncigar.p, 0, b->n * 4, st));
	CK(cudaMemsetAsync(b->d_dense_total.p, 0, 16, st));
	CK(cudaMemsetAsync(b->d_dense_off.p, 0, b->n * 8, st));
	// four events per wave: [forward start, forward done, traceback start, traceback done]
	evs.v.assign(b->waves.size() * 4, nullptr);
	for(auto &e : evs.v) CK(cudaEventCreate(&e));
	// the arena may have been trimmed or re-planned smaller since this batch was uploaded (bsb200_trim, another batch's allocation failure)
	if(ctx->trace.cap < b->max_wave_bytes + 256){
		if(ctx->trace.reserve(b->max_wave_bytes + 256, true) != cudaSuccess){ cudaGetLastError(); return fail(ctx, "traceback arena smaller than this batch's largest wave and it cannot be re-allocated", cudaErrorMemoryAllocation); }
	}
	cudaStream_t sb = st; // (ctx->stream_bt is kept for experiments with concurrent traceback)
	for(size_t wi=0;wi<b->waves.size();wi++){
		const Wave &w = b->waves[wi];
		uint32_t np = w.end - w.beg;
		uint8_t *arena = ctx->trace.as<uint8_t>();
		CK(cudaMemsetAsync(ctx->counter.p, 0, 16, st));
		CK(cudaEventRecord(evs[wi * 4 + 0], st));
		if(b->kind == 0){
			Epi8Args a;
			a.seqs = b->seqs_dev(); a.qoff = b->d_qoff.as<uint64_t>(); a.toff = b->d_toff.as<uint64_t>();
			a.qlen = b->d_qlen.as<uint32_t>(); a.tlen = b->d_tlen.as<uint32_t>();
			a.order = b->d_order.as<uint32_t>() + w.beg; a.npairs = np; a.counter = ctx->counter.as<unsigned int>();
			a.trace = arena; a.trace_off = b->d_trace_off.as<uint64_t>();
			a.results = b->d_results.as<int32_t>(); a.status = b->d_status.as<int32_t>();
			a.bandwidth = b->bandwidth; a.max_img = epi8_image_bytes(b->max_bw / 16);
			// images of u,(e),(q) and the selectors + anchors and scratch; 32 bytes past a multiple of 128 so the four
			// groups of a warp keep their small per-group words (anchors, F hand-over) on different banks
			a.group_smem = (uint32_t)(((size_t)a.max_img * (b->pw + 2) + (kMetaInts * 2) * 4 + 32 + 32 * 4 + 127) / 128 * 128 + 32);
			a.mode = b->mode; memcpy(a.mtx, b->mtx, 16); a.go1 = b->go1; a.ge1 = b->ge1; a.go2 = b->go2; a.ge2 = b->ge2;
			a.all_ones = 0xffffffffu; a.c256 = 256u; a.c65536 = 65536u; a.redo = 0; a.force_redo = 0;
			a.smax = -127; a.smin = 127;
			for(int k=0;k<16;k++){ a.smax = std::max(a.smax, b->mtx[k]); a.smin = std::min(a.smin, b->mtx[k]); }
			// all gap costs <= 0 (the normal case): saturation bounds that cannot bind are dropped (epi8_forward.cuh)
			// (linear gaps, pw = 0, stay on the literal kernel)
			// (BSB200_NOFAST: tests run the literal kernels on ordinary gap costs too)
			const bool fast = b->pw >= 1 && b->ge1 <= 0 && (int8_t)(b->go1 + b->ge1) <= 0 && (b->pw < 2 || (b->ge2 <= 0 && (int8_t)(b->go2 + b->ge2) <= 0)) && !getenv("BSB200_NOFAST");
			const bool anch = b->wave_split ? epi8_wave_use_anchors(b->max_bw / 16) : epi8_use_anchors(b->max_bw / 16);
			a.gpw = 4;
			int rc;
			if(b->wave_split){
				// the single-pass wavefront kernel (epi8_wave.cuh), then the two-pass kernel below over the pairs it flagged
				// (normally none: that launch only reads the status words)
				Epi8Args wa = a;
				wa.group_smem = (uint32_t)wave_group_smem(b->max_bw, b->wave_split);
				wa.force_redo = getenv("BSB200_WAVE_REDO") ? 1 : 0;
				rc = launch_epi8_wave(ctx, wa, np, anch, b->wave_split);
				if(rc) return rc;
				ctx->timing.forward_launches++;
				CK(cudaMemsetAsync(ctx->counter.p, 0, 16, st));
				a.redo = 1;
			}
			// every band covers its whole query: bandwidth 0, or a bandwidth (rounded up to 16 by the kernel) no shorter than the longest query
			const bool full = (b->bandwidth == 0 || (b->bandwidth + 15) / 16 * 16 >= b->max_qlen) && !getenv("BSB200_NOFULL");   // (BSB200_NOFULL: tests run the general instantiation on full bands too)
			a.bulk_store = bulk_store_default(full);
			if(fast) rc = anch ? launch_epi8_forward_pw<true, true>(ctx, a, np, b->pw, full) : launch_epi8_forward_pw<true, false>(ctx, a, np, b->pw, full);
			else rc = anch ? launch_epi8_forward_pw<false, true>(ctx, a, np, b->pw, full) : launch_epi8_forward_pw<false, false>(ctx, a, np, b->pw, full);
			if(rc) return rc;
			ctx->timing.forward_launches++;
			CK(cudaEventRecord(evs[wi * 4 + 1], st));
			// traceback of this wave on the second stream, behind its forward kernel
			CK(cudaEventRecord(evs[wi * 4 + 2], sb));
			Epi8BtArgs t;
			t.seqs = a.seqs; t.qoff = a.qoff; t.toff = a.toff; t.qlen = a.qlen; t.tlen = a.tlen; t.order = a.order; t.npairs = np;
			t.trace = a.trace; t.trace_off = a.trace_off; t.results = a.results; t.status = a.status;
			t.cigars = b->want_cigar ? b->d_cig_raw.as<uint32_t>() : nullptr; t.cig_off = b->d_cig_off.as<uint64_t>();
			t.dense = b->want_cigar ? b->d_cig_dense.as<uint32_t>() : nullptr; t.dense_off = b->d_dense_off.as<uint64_t>();
			t.dense_total = b->d_dense_total.as<unsigned long long>();
			t.ncigar = b->d_ncigar.as<uint32_t>();
			t.bandwidth = b->bandwidth; t.mode = b->mode; t.pw = b->pw; t.split = b->wave_split; t.ubias = fast ? 128 : 0; t.anch = anch ? 1 : 0; memcpy(t.mtx, b->mtx, 16);
			t.go1 = b->go1; t.ge1 = b->ge1; t.go2 = b->go2; t.ge2 = b->ge2;
			// one walk per thread; a batch too small to fill the GPU with warps spreads its walks (one every `stride` threads)
			{
				// (as long as every block is resident at once: a second round of blocks would double the time)
				int per_sm = 1;
				cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, epi8_backcal_kernel, 64, 0);
				const uint64_t resident = (uint64_t)std::max(per_sm, 1) * ctx->num_sms * 64;
				uint32_t stride = 1;
				while(stride < 32 && (uint64_t)np * stride * 2 <= resident) stride *= 2;
				if(const char *ev = getenv("BSB200_BT_S")){ const int v = atoi(ev); if(v == 1 || v == 2 || v == 4 || v == 8 || v == 16 || v == 32) stride = (uint32_t)v; }
				t.stride = stride;
				const uint64_t threads = (uint64_t)np * stride;
				epi8_backcal_kernel<<<(unsigned)((threads + 63) / 64), 64, 0, sb>>>(t);
			}
			CK(cudaGetLastError());
			ctx->timing.traceback_launches++;
			CK(cudaEventRecord(evs[wi * 4 + 3], sb));
		} else {
			EditArgs a;
			a.seqs = b->seqs_dev(); a.qoff = b->d_qoff.as<uint64_t>(); a.toff = b->d_toff.as<uint64_t>();
			a.qlen = b->d_qlen.as<uint32_t>(); a.tlen = b->d_tlen.as<uint32_t>();
			a.order = b->d_order.as<uint32_t>() + w.beg; a.npairs = np;
			a.trace = ctx->trace.as<uint8_t>();
			a.block_off = b->d_trace_off.as<uint64_t>() + w.beg / 32; a.block_rows = b->d_block_rows.as<uint32_t>() + w.beg / 32;
			a.results = b->d_results.as<int32_t>(); a.status = b->d_status.as<int32_t>();
			a.cigars = b->want_cigar ? b->d_cig_raw.as<uint32_t>() : nullptr; a.cig_off = b->d_cig_off.as<uint64_t>();
			a.dense = b->want_cigar ? b->d_cig_dense.as<uint32_t>() : nullptr; a.dense_off = b->d_dense_off.as<uint64_t>();
			a.dense_total = b->d_dense_total.as<unsigned long long>(); a.ncigar = b->d_ncigar.as<uint32_t>();
			a.mode = b->mode; a.bandwidth = b->bandwidth; a.WB = b->max_bw / 64; a.nQW = b->max_q64 / 64 + 2;
			int rc = launch_edit(ctx, a);
			if(rc) return rc;
			ctx->timing.forward_launches++;
			CK(cudaEventRecord(evs[wi * 4 + 1], st));
			CK(cudaEventRecord(evs[wi * 4 + 2], st));
			CK(cudaEventRecord(evs[wi * 4 + 3], st));
		}
	}
	if(!b->long_pairs.empty()){
		// edit pairs beyond edit_kernel's limits: one thread per pair on scratch in HBM (edit_long_kernel), as many pairs at a time as 16 GB hold
		const uint64_t cap = 16ull << 30;
		DevBuf &scr = ctx->kmer_cache[1], &ids = ctx->kmer_cache[3], &offs = ctx->kmer_cache[4];
		for(size_t k0=0;k0<b->long_pairs.size();){
			size_t k1 = k0 + 1;
			const uint64_t base = b->long_off[k0];
			auto end_of = [&](size_t k){ return k + 1 < b->long_pairs.size() ? b->long_off[k + 1] : b->long_bytes; };
			while(k1 < b->long_pairs.size() && end_of(k1) - base <= cap) k1++;
			const uint64_t bytes = end_of(k1 - 1) - base;
			const size_t m = k1 - k0;
			if(scr.reserve(bytes + 256, true) != cudaSuccess){ cudaGetLastError(); ctx->err = "edit: not enough device memory for the trace of a pair longer than the kernel's band limit"; return -1; }
			CK(ids.reserve(m * 4 + 16)); CK(offs.reserve(m * 8 + 16));
			std::vector<uint64_t> rel(m);
			for(size_t k=0;k<m;k++) rel[k] = b->long_off[k0 + k] - base;
			CK(cudaMemcpyAsync(ids.p, b->long_pairs.data() + k0, m * 4, cudaMemcpyHostToDevice, st));
			CK(cudaMemcpyAsync(offs.p, rel.data(), m * 8, cudaMemcpyHostToDevice, st));
			EditLongArgs a;
			a.seqs = b->seqs_dev(); a.qoff = b->d_qoff.as<uint64_t>(); a.toff = b->d_toff.as<uint64_t>(); a.qlen = b->d_qlen.as<uint32_t>(); a.tlen = b->d_tlen.as<uint32_t>();
			a.pairs = ids.as<uint32_t>(); a.scr_off = offs.as<uint64_t>(); a.npairs = (uint32_t)m; a.scratch = scr.as<uint8_t>(); a.mode = b->mode;
			a.results = b->d_results.as<int32_t>(); a.status = b->d_status.as<int32_t>();
			a.cigars = b->want_cigar ? b->d_cig_raw.as<uint32_t>() : nullptr; a.cig_off = b->d_cig_off.as<uint64_t>();
			a.dense = b->want_cigar ? b->d_cig_dense.as<uint32_t>() : nullptr; a.dense_off = b->d_dense_off.as<uint64_t>();
			a.dense_total = b->d_dense_total.as<unsigned long long>(); a.ncigar = b->d_ncigar.as<uint32_t>();
			edit_long_kernel<<<(unsigned)m, 32, 0, st>>>(a);
			CK(cudaGetLastError());
			CK(cudaStreamSynchronize(st));   // rel is a local
			ctx->timing.other_launches++;
			k0 = k1;
		}
	}
	CK(cudaEventRecord(ctx->ev[5], st));
	CK(cudaStreamSynchronize(st));
	{ float rm = 0; cudaEventElapsedTime(&rm, ctx->ev[4], ctx->ev[5]); ctx->timing.run_ms = rm; }
	for(size_t wi=0;wi<b->waves.size();wi++){
		float m1 = 0, m2 = 0;
		cudaEventElapsedTime(&m1, evs[wi * 4 + 0], evs[wi * 4 + 1]);
		cudaEventElapsedTime(&m2, evs[wi * 4 + 2], evs[wi * 4 + 3]);
		fwd_ms += m1; bt_ms += m2;
	}
	ctx->timing.forward_ms = fwd_ms; ctx->timing.traceback_ms = bt_ms;
	ctx->timing.other_launches += 5 + (uint32_t)b->waves.size();
	b->ran = true;
	return 0;
}

extern "C" int bsb200_batch_sync(bsb200_ctx *ctx){
	if(!ctx) return -1;
	CK(cudaStreamSynchronize(ctx->stream));
	return 0;
}

// ---- dense cigars in PAIR ORDER: prefix sum of the per-pair counts, then one warp per pair moves its words ---------
struct U32toU64 { __host__ __device__ uint64_t operator()(const uint32_t &x) const { return (uint64_t)x; } };

__global__ void __launch_bounds__(256) cigar_order_kernel(const uint32_t *dense, const uint64_t *dense_off, const uint32_t *ncigar, const uint64_t *prefix, uint32_t *ordered, uint64_t n){
	uint64_t pair = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
	if(pair >= n) return;
	const uint32_t lane = threadIdx.x & 31, k = ncigar[pair];
	const uint32_t *src = dense + dense_off[pair];
	uint32_t *dst = ordered + prefix[pair];
	for(uint32_t i=lane;i<k;i+=32) dst[i] = src[i];
}

static int fetch_impl(bsb200_ctx *ctx, bsb200_batch *b, bsb200_result_t *results, uint32_t *cigars, const uint64_t *cgoff,
		uint64_t dense_cap, uint64_t *total_out, uint32_t *ncigar, int32_t *status);

extern "C" int bsb200_batch_fetch(bsb200_ctx *ctx, bsb200_batch *b, bsb200_result_t *results,
		uint32_t *cigars, const uint64_t *cgoff, uint32_t *ncigar, int32_t *status){
	return fetch_impl(ctx, b, results, cigars, cgoff, 0, nullptr, ncigar, status);
}

extern "C" int bsb200_batch_fetch_dense(bsb200_ctx *ctx, bsb200_batch *b, bsb200_result_t *results,
		uint32_t *cigars, uint64_t cigar_cap_words, uint64_t *total_words, uint32_t *ncigar, int32_t *status){
	return fetch_impl(ctx, b, results, cigars, nullptr, cigar_cap_words, total_words, ncigar, status);
}

static int fetch_impl(bsb200_ctx *ctx, bsb200_batch *b, bsb200_result_t *results,
		uint32_t *cigars, const uint64_t *cgoff, uint64_t dense_cap, uint64_t *total_out, uint32_t *ncigar, int32_t *status){
	if(!ctx || !b || !b->ran) return fail(ctx, "bsb200_batch_fetch before run", cudaSuccess);
	ctx->err.clear();
	cudaSetDevice(ctx->device);
	cudaStream_t st = ctx->stream;
	uint64_t n = b->n;
	if(n == 0) return 0;
	const bool cg = b->want_cigar && cigars && cgoff;
	const bool cgd = b->want_cigar && cigars && !cgoff;   // dense, pair-ordered output straight into the caller's buffer
	if(total_out) *total_out = 0;
	CK(b->h_results.reserve(n * 40)); CK(b->h_status.reserve(n * 4)); CK(b->h_ncigar.reserve(n * 4)); CK(b->h_total.reserve(16));
	cudaEventRecord(ctx->ev[2], st);
	CK(d2h_copy(ctx, results ? (void*)results : b->h_results.p, b->d_results.p, n * 40, st));
	CK(cudaMemcpyAsync(b->h_status.p, b->d_status.p, n * 4, cudaMemcpyDeviceToHost, st));
	CK(cudaMemcpyAsync(b->h_ncigar.p, b->d_ncigar.p, n * 4, cudaMemcpyDeviceToHost, st));
	uint64_t d2h = n * 48;
	uint64_t total = 0;
	if(cg){
		CK(b->h_dense_off.reserve(n * 8));
		CK(cudaMemcpyAsync(b->h_dense_off.p, b->d_dense_off.p, n * 8, cudaMemcpyDeviceToHost, st));
		CK(cudaMemcpyAsync(b->h_total.p, b->d_dense_total.p, 8, cudaMemcpyDeviceToHost, st));
		CK(cudaStreamSynchronize(st));
		total = *b->h_total.as<unsigned long long>();
		CK(b->h_dense.reserve(total * 4 + 16));
		if(total) CK(cudaMemcpyAsync(b->h_dense.p, b->d_cig_dense.p, total * 4, cudaMemcpyDeviceToHost, st));
		d2h += n * 8 + 8 + total * 4;
	}
	if(cgd){
		// prefix[i] = words of pairs < i (device scan), one warp per pair copies into pair order, one D2H of the result
		CK(b->d_prefix.reserve((n + 1) * 8));
		size_t tmp_bytes = 0;
		cub::TransformInputIterator<uint64_t, U32toU64, const uint32_t*> it(b->d_ncigar.as<uint32_t>(), U32toU64());
		CK(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, it, b->d_prefix.as<uint64_t>(), (int)n, st));
		CK(ctx->counter.reserve(tmp_bytes + 256));
		CK(cub::DeviceScan::ExclusiveSum((uint8_t*)ctx->counter.p + 256, tmp_bytes, it, b->d_prefix.as<uint64_t>(), (int)n, st));
		CK(cudaMemcpyAsync(b->h_total.p, b->d_dense_total.p, 8, cudaMemcpyDeviceToHost, st));
		CK(cudaStreamSynchronize(st));
		total = *b->h_total.as<unsigned long long>();
		if(total_out) *total_out = total;
		if(total > dense_cap) return fail(ctx, "cigar buffer too small for the dense cigars (see total_words)", cudaSuccess);
		if(total){
			uint32_t *ordered = b->d_cig_raw.as<uint32_t>(); // the per-pair scratch regions are free again once the walks are done
			cigar_order_kernel<<<(unsigned)((n * 32 + 255) / 256), 256, 0, st>>>(b->d_cig_dense.as<uint32_t>(), b->d_dense_off.as<uint64_t>(),
				b->d_ncigar.as<uint32_t>(), b->d_prefix.as<uint64_t>(), ordered, n);
			CK(cudaGetLastError());
			CK(d2h_copy(ctx, cigars, ordered, total * 4, st));
		}
		d2h += 8 + total * 4;
	}
	cudaEventRecord(ctx->ev[3], st);
	CK(cudaStreamSynchronize(st));
	float ms = 0; cudaEventElapsedTime(&ms, ctx->ev[2], ctx->ev[3]);
	ctx->timing.d2h_ms = ms; ctx->timing.d2h_bytes = d2h;
	{ int32_t *hs0 = b->h_status.as<int32_t>(); for(uint64_t i=0;i<n;i++) if(b->empty[i] == 1) hs0[i] |= BSB200_ST_EMPTY; else if(b->empty[i] == 3) hs0[i] |= BSB200_ST_UNSUPPORTED; }
	if(status) memcpy(status, b->h_status.p, n * 4);
	const uint32_t *hn = b->h_ncigar.as<uint32_t>();
	if(ncigar) memcpy(ncigar, hn, n * 4);
	if(cg){
		const uint64_t *doff = b->h_dense_off.as<uint64_t>();
		const uint32_t *dense = b->h_dense.as<uint32_t>();
		int32_t *hs = b->h_status.as<int32_t>();
		for(uint64_t i=0;i<n;i++){
			uint64_t cap = cgoff[i + 1] - cgoff[i];
			uint64_t k = hn[i];
			if(k > cap){ k = cap; if(status) status[i] |= BSB200_ST_CIGCAP; hs[i] |= BSB200_ST_CIGCAP; }
			if(k) memcpy(cigars + cgoff[i], dense + doff[i], k * 4);
		}
	}
	return 0;
}

// Results of a finished run into DEVICE buffers of the caller (they travel on to another GPU over NVLink: bsalign_b200/shard.py):
// d_results n x 10 int32, d_ncigar n, d_status n (kernel flags; BSB200_ST_EMPTY is a host-side flag - pairs with qlen or tlen 0
// come back all zero), d_cigars dense and pair-ordered like bsb200_batch_fetch_dense.  *total_words always receives the word count.
extern "C" int bsb200_batch_fetch_dense_dev(bsb200_ctx *ctx, bsb200_batch *b, int32_t *d_results, uint32_t *d_cigars, uint64_t cigar_cap_words,
		uint64_t *total_words, uint32_t *d_ncigar, int32_t *d_status){
	if(!ctx || !b || !b->ran) return fail(ctx, "bsb200_batch_fetch_dense_dev before run", cudaSuccess);
	ctx->err.clear();
	cudaSetDevice(ctx->device);
	cudaStream_t st = ctx->stream;
	const uint64_t n = b->n;
	if(total_words) *total_words = 0;
	if(n == 0) return 0;
	cudaEventRecord(ctx->ev[2], st);
	if(d_results) CK(cudaMemcpyAsync(d_results, b->d_results.p, n * 40, cudaMemcpyDeviceToDevice, st));
	if(d_status) CK(cudaMemcpyAsync(d_status, b->d_status.p, n * 4, cudaMemcpyDeviceToDevice, st));
	if(d_ncigar) CK(cudaMemcpyAsync(d_ncigar, b->d_ncigar.p, n * 4, cudaMemcpyDeviceToDevice, st));
	uint64_t total = 0;
	if(b->want_cigar){
		CK(b->h_total.reserve(16));
		CK(b->d_prefix.reserve((n + 1) * 8));
		size_t tmp_bytes = 0;
		cub::TransformInputIterator<uint64_t, U32toU64, const uint32_t*> it(b->d_ncigar.as<uint32_t>(), U32toU64());
		CK(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, it, b->d_prefix.as<uint64_t>(), (int)n, st));
		CK(ctx->counter.reserve(tmp_bytes + 256));
		CK(cub::DeviceScan::ExclusiveSum((uint8_t*)ctx->counter.p + 256, tmp_bytes, it, b->d_prefix.as<uint64_t>(), (int)n, st));
		CK(cudaMemcpyAsync(b->h_total.p, b->d_dense_total.p, 8, cudaMemcpyDeviceToHost, st));
		CK(cudaStreamSynchronize(st));
		total = *b->h_total.as<unsigned long long>();
		if(total_words) *total_words = total;
		if(d_cigars){
			if(total > cigar_cap_words) return fail(ctx, "cigar buffer too small for the dense cigars (see total_words)", cudaSuccess);
			if(total){
				cigar_order_kernel<<<(unsigned)((n * 32 + 255) / 256), 256, 0, st>>>(b->d_cig_dense.as<uint32_t>(), b->d_dense_off.as<uint64_t>(),
					b->d_ncigar.as<uint32_t>(), b->d_prefix.as<uint64_t>(), d_cigars, n);
				CK(cudaGetLastError());
			}
		}
	}
	cudaEventRecord(ctx->ev[3], st);
	CK(cudaStreamSynchronize(st));
	float ms = 0; cudaEventElapsedTime(&ms, ctx->ev[2], ctx->ev[3]);
	ctx->timing.d2h_ms = ms; ctx->timing.d2h_bytes = 8;
	return 0;
}


// ---- k-mer guided edit alignment: replaces kmer_striped_seqedit_pairwise (bsalign.h:1209) for batches ------------------------------
// Anchors, chain, filter, the gap alignments and the stitching run in kmer_edit_kernel (kmer_edit.cuh), one warp per pair.  Pairs without
// usable anchors come back flagged: the reference aligns those with the plain global edit (bsalign.h:1440), and so do we, as a sub-batch
// of the edit kernel on the arena that is already on the device.  Output as in bsb200_batch_fetch (cgoff given) or bsb200_batch_fetch_dense.
static int dense_pipelined(bsb200_ctx *ctx, int kind, uint64_t n, const uint8_t *seqs, const uint64_t *bits,
		const uint64_t *qoff, const uint32_t *qlen, const uint64_t *toff, const uint32_t *tlen,
		int mode, uint32_t bandwidth, const int8_t *matrix, int8_t go1, int8_t ge1, int8_t go2, int8_t ge2,
		bsb200_result_t *results, uint32_t *cigars, uint64_t cigar_cap_words, uint64_t *total_words, uint32_t *ncigar, int32_t *status, uint32_t ksz);

// a chunk of a pipelined call fetches its dense cigars behind those of the chunks before it: begin() waits for the turn and returns the
// words already out, end() publishes this chunk's count
struct FetchTurn { std::function<uint64_t()> begin; std::function<void(uint64_t)> end; bool taken = false; };

static int kmer_edit_impl(bsb200_ctx *ctx, uint64_t n, const uint8_t *seqs, const uint64_t *qoff, const uint32_t *qlen, const uint64_t *toff,
		const uint32_t *tlen, uint32_t ksz, bsb200_result_t *results, uint32_t *cigars, const uint64_t *cgoff, uint64_t dense_cap, uint64_t *total_words,
		uint32_t *ncigar, int32_t *status, const SeqSegs *segs = nullptr, FetchTurn *turn = nullptr, int share = 1){
	if(!ctx) return -1;
	ctx->err.clear();
	if(total_words) *total_words = 0;
	if(ksz > 15) ksz = 15;   // bsalign.h:1217
	if(ksz == 0 || n >= 0xFFFFFFF0ull || (n && (!seqs || !qoff || !qlen || !toff || !tlen))) return fail(ctx, "bsb200_kmer_edit_batch (k-mer size 1..15)", cudaSuccess);
	if(n == 0) return 0;
	cudaSetDevice(ctx->device);
	cudaStream_t st = ctx->stream;
	const bool hostprof = getenv("BSB200_HOSTPROF") != nullptr;
	auto tp0 = std::chrono::steady_clock::now();
	auto lap = [&](const char *what){ if(hostprof){ auto t1 = std::chrono::steady_clock::now(); fprintf(stderr, "[kmer] %-28s %8.3f ms\n", what, std::chrono::duration<double, std::milli>(t1 - tp0).count()); tp0 = t1; } };
	bsb200_batch *b = new bsb200_batch();
	{
		DevBuf *ds[] = {&b->d_seqs, &b->d_qoff, &b->d_toff, &b->d_qlen, &b->d_tlen, &b->d_order, &b->d_trace_off, &b->d_results, &b->d_status,
			&b->d_ncigar, &b->d_cig_raw, &b->d_cig_off, &b->d_cig_dense, &b->d_dense_off, &b->d_dense_total, &b->d_block_rows, &b->d_prefix, &b->d_bits};
		for(int k=0;k<18;k++){ *ds[k] = ctx->dev_cache[k]; ctx->dev_cache[k] = DevBuf(); }
		HostBuf *hs[] = {&b->h_results, &b->h_status, &b->h_ncigar, &b->h_dense_off, &b->h_dense, &b->h_total};
		for(int k=0;k<6;k++){ *hs[k] = ctx->host_cache[k]; ctx->host_cache[k] = HostBuf(); }
	}
	b->kind = 1; b->n = n; b->mode = 0; b->bandwidth = 0; b->want_cigar = 1;
	cudaEvent_t evk[2] = {nullptr, nullptr};
	auto done = [&](int rc){ cudaStreamSynchronize(st); for(auto &e_ : evk) if(e_) cudaEventDestroy(e_); bsb200_batch_free(ctx, b); return rc; };
	#define CKK(call) do { cudaError_t _e = (call); if(_e != cudaSuccess) return done(fail(ctx, #call, _e)); } while(0)
	const int NT = plan_threads(n);
	size_t seq_end = 0; uint32_t maxq = 0, maxt = 0; uint64_t cells = 0;
	b->empty.assign(n, 0);
	{
		std::vector<uint64_t> pe(NT, 0), pc(NT, 0), px(NT, 0); std::vector<uint32_t> mq(NT, 0), mt(NT, 0);
		par_slices(n, NT, [&](int w, uint64_t lo, uint64_t hi){
			uint64_t se = 0, cc = 0, cx = 0; uint32_t q_ = 0, t_ = 0;   // locals: the per-thread slots share cache lines
			for(uint64_t i=lo;i<hi;i++){
				se = std::max<uint64_t>(se, std::max<uint64_t>(qoff[i] + qlen[i], toff[i] + tlen[i]));
				if(qlen[i] && tlen[i]){ cc += (uint64_t)qlen[i] + tlen[i] + 2; cx += (uint64_t)qlen[i] * tlen[i]; q_ = std::max(q_, qlen[i]); t_ = std::max(t_, tlen[i]); }
				else b->empty[i] = 1;
			}
			pe[w] = se; pc[w] = cc; px[w] = cx; mq[w] = q_; mt[w] = t_;
		});
		for(int w=0;w<NT;w++){ seq_end = std::max<size_t>(seq_end, pe[w]); b->cig_words += pc[w]; cells += px[w]; maxq = std::max(maxq, mq[w]); maxt = std::max(maxt, mt[w]); }
	}
	lap("scan of the lengths");
	CKK(b->d_seqs.reserve(seq_end + 16));
	CKK(b->d_qoff.reserve(n * 8 + 8)); CKK(b->d_toff.reserve(n * 8 + 8)); CKK(b->d_qlen.reserve(n * 4 + 4)); CKK(b->d_tlen.reserve(n * 4 + 4));
	lap("input buffers");
	CKK(b->d_results.reserve(n * 40 + 40)); CKK(b->d_status.reserve(n * 4 + 4)); CKK(b->d_ncigar.reserve(n * 4 + 4));
	CKK(b->d_dense_off.reserve(n * 8 + 8)); CKK(b->d_dense_total.reserve(16));
	CKK(b->d_cig_raw.reserve(b->cig_words * 4 + 16)); CKK(b->d_cig_dense.reserve(b->cig_words * 4 + 16));
	CKK(cudaEventCreate(&evk[0])); CKK(cudaEventCreate(&evk[1]));
	lap("output buffers + events");
	cudaEventRecord(ctx->ev[0], st);
	if(segs){
		CKK(h2d_copy(ctx, b->d_seqs.p, seqs, segs->len0, st));
		CKK(h2d_copy(ctx, b->d_seqs.as<uint8_t>() + segs->at1, seqs + segs->src1, segs->len1, st));
	} else CKK(h2d_copy(ctx, b->d_seqs.p, seqs, seq_end, st));
	CKK(cudaMemcpyAsync(b->d_qoff.p, qoff, n * 8, cudaMemcpyHostToDevice, st));
	CKK(cudaMemcpyAsync(b->d_toff.p, toff, n * 8, cudaMemcpyHostToDevice, st));
	CKK(cudaMemcpyAsync(b->d_qlen.p, qlen, n * 4, cudaMemcpyHostToDevice, st));
	CKK(cudaMemcpyAsync(b->d_tlen.p, tlen, n * 4, cudaMemcpyHostToDevice, st));
	cudaEventRecord(ctx->ev[1], st);
	// ---- warp slots and the pool for large gap traces ------------------------------------------------------------------------
	KmerArgs a;
	memset(&a, 0, sizeof(a));
	size_t freeb = 0, totalb = 0;
	CKK(cudaMemGetInfo(&freeb, &totalb));
	freeb += ctx->kmer_cache[0].cap + ctx->kmer_cache[1].cap;
	// hash-table keys of short pairs in shared memory: up to 2048 entries per warp (32 KB per CTA) keep seven of the eight CTAs of an SM resident
	uint64_t Hmax = 64; while(2 * Hmax < 3 * ((uint64_t)maxq + maxt)) Hmax <<= 1;
	const uint32_t smem_keys = (Hmax <= 2048 && !getenv("BSB200_KMER_NOSMEM")) ? (uint32_t)Hmax : 0;
	int occ = 0;
	CKK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kmer_edit_kernel, kKmWarps * 32, (size_t)smem_keys * 4 * kKmWarps));
	if(occ < 1) occ = 1;
	const uint64_t max_warps = std::max<uint64_t>(kKmWarps, (uint64_t)ctx->num_sms * occ * kKmWarps / (share > 1 ? share : 1));   // share: calls running side by side on this device
	// pairs a warp works on at a time: as many as leave every warp of the GPU at least two groups (a group's serial phases use one lane per pair)
	uint32_t group = 1;
	while(group < 32 && n >= 2 * max_warps * (group * 2)) group *= 2;
	if(const char *e_ = getenv("BSB200_KMER_GROUP")){ const long g_ = strtol(e_, nullptr, 10); group = (uint32_t)std::min<long>(32, std::max<long>(1, g_)); }
	uint64_t slots = std::min<uint64_t>(max_warps, ((n + group - 1) / group + kKmWarps - 1) / kKmWarps * kKmWarps);
	if(const char *e_ = getenv("BSB200_KMER_SLOTS")) slots = std::max<uint64_t>(kKmWarps, strtoull(e_, nullptr, 10) / kKmWarps * kKmWarps);
	uint64_t wb = kmer_warp_bytes(maxq, maxt, group, &a);
	while(group > 1 && slots * wb > freeb / 2){ group /= 2; wb = kmer_warp_bytes(maxq, maxt, group, &a); }
	while(slots > kKmWarps && slots * wb > freeb / 2) slots = (slots / 2 + kKmWarps - 1) / kKmWarps * kKmWarps;
	if(slots * wb > freeb / 2) return done(fail(ctx, "k-mer edit: a pair of this length does not fit the device scratch", cudaSuccess));
	uint64_t pool_bytes = std::min<uint64_t>((freeb - slots * wb) / 2, 16ull << 30);
	if(const char *e_ = getenv("BSB200_KMER_POOL")) pool_bytes = strtoull(e_, nullptr, 10);
	pool_bytes = std::max<uint64_t>(pool_bytes, 4096) & ~15ull;
	CKK(ctx->kmer_cache[0].reserve(slots * wb, true));
	CKK(ctx->kmer_cache[1].reserve(pool_bytes, true));
	CKK(ctx->kmer_cache[2].reserve(256));
	a.seqs = b->d_seqs.as<uint8_t>(); a.qoff = b->d_qoff.as<uint64_t>(); a.toff = b->d_toff.as<uint64_t>();
	a.qlen = b->d_qlen.as<uint32_t>(); a.tlen = b->d_tlen.as<uint32_t>();
	a.ksz = ksz;
	a.smem_keys = smem_keys;
	a.next = ctx->kmer_cache[2].as<unsigned int>();
	a.pool_used = (unsigned long long*)((uint8_t*)ctx->kmer_cache[2].p + 16);
	a.scratch = ctx->kmer_cache[0].as<uint8_t>();
	a.pool = ctx->kmer_cache[1].as<uint8_t>(); a.pool_bytes = pool_bytes;
	a.results = b->d_results.as<int32_t>(); a.status = b->d_status.as<int32_t>(); a.ncigar = b->d_ncigar.as<uint32_t>();
	a.dense = b->d_cig_dense.as<uint32_t>(); a.dense_off = b->d_dense_off.as<uint64_t>(); a.dense_total = b->d_dense_total.as<unsigned long long>();
	CKK(cudaMemsetAsync(b->d_dense_total.p, 0, 16, st));
	CKK(b->h_status.reserve(n * 4));
	lap("copies queued, scratch");
	int32_t *hst = b->h_status.as<int32_t>();
	std::vector<uint32_t> redo, fall;
	uint32_t launches = 0;
	cudaEventRecord(evk[0], st);
	for(int round=0;;round++){
		// round 0: every pair.  Round 1: the pairs that found the pool empty, with two CTAs sharing it.  Round 2: what is left, ONE pair
		// per launch (the pool is handed out by a bump pointer and only the host resets it).
		const bool single = round >= 2;
		const uint64_t np = round == 0 ? n : redo.size();
		if(round > 0){
			CKK(ctx->kmer_cache[3].reserve(redo.size() * 4 + 16));
			CKK(cudaMemcpyAsync(ctx->kmer_cache[3].p, redo.data(), redo.size() * 4, cudaMemcpyHostToDevice, st));
		}
		for(uint64_t k=0;k<(single ? np : 1);k++){
			a.order = round == 0 ? nullptr : ctx->kmer_cache[3].as<uint32_t>() + (single ? k : 0);
			a.npairs = single ? 1u : (uint32_t)np;
			if(single) a.group = 1;
			CKK(cudaMemsetAsync(ctx->kmer_cache[2].p, 0, 32, st));
			const uint64_t warps = round == 0 ? slots : std::min<uint64_t>(slots, 2 * kKmWarps);
			kmer_edit_kernel<<<single ? 1u : (unsigned)((warps + kKmWarps - 1) / kKmWarps), single ? 32 : kKmWarps * 32, (size_t)a.smem_keys * 4 * kKmWarps, st>>>(a);
			CKK(cudaGetLastError());
			launches++;
		}
		CKK(cudaMemcpyAsync(hst, b->d_status.p, n * 4, cudaMemcpyDeviceToHost, st));
		CKK(cudaStreamSynchronize(st));
		std::vector<uint32_t> again;
		if(round == 0){ for(uint64_t i=0;i<n;i++){ if(hst[i] & kStPool) again.push_back((uint32_t)i); else if(hst[i] & kStFallback) fall.push_back((uint32_t)i); } }
		else for(uint32_t i : redo){ if(hst[i] & kStPool) again.push_back(i); else if(hst[i] & kStFallback) fall.push_back(i); }
		redo.swap(again);
		if(redo.empty()) break;
		if(single) return done(fail(ctx, "k-mer edit: the gap traces of one pair are larger than the device pool", cudaSuccess));
	}
	cudaEventRecord(evk[1], st);
	lap("kernel rounds");
	// ---- pairs without anchors: plain global edit (bsalign.h:1440) on the resident arena, results merged back -------------------------
	float fb_ms = 0; uint32_t fb_launches = 0;
	if(!fall.empty()){
		const uint64_t m = fall.size();
		std::sort(fall.begin(), fall.end());
		std::vector<uint64_t> sq(m), so(m); std::vector<uint32_t> ql(m), tl(m);
		for(uint64_t f=0;f<m;f++){ const uint32_t i = fall[f]; sq[f] = qoff[i]; so[f] = toff[i]; ql[f] = qlen[i]; tl[f] = tlen[i]; }
		bsb200_batch *sb = upload_impl(ctx, 1, m, nullptr, b->d_seqs.as<uint8_t>(), nullptr, sq.data(), ql.data(), so.data(), tl.data(), 0, 0, nullptr, 0, 0, 0, 0, 1);
		if(!sb) return done(-1);
		int rc = bsb200_batch_run(ctx, sb);
		fb_ms = ctx->timing.run_ms; fb_launches = ctx->timing.forward_launches + ctx->timing.traceback_launches + ctx->timing.other_launches;
		cudaError_t e1 = cudaSuccess;
		if(rc == 0){
			e1 = ctx->kmer_cache[4].reserve(m * 40 + 40);
			if(e1 == cudaSuccess) e1 = ctx->kmer_cache[5].reserve(m * 4 + 4);
			if(e1 == cudaSuccess) e1 = ctx->kmer_cache[6].reserve(m * 4 + 4);
			if(e1 == cudaSuccess) e1 = ctx->kmer_cache[7].reserve(sb->cig_words * 4 + 16);
			if(e1 != cudaSuccess){ fail(ctx, "device allocation", e1); rc = -1; }
		}
		uint64_t tw = 0;
		if(rc == 0) rc = bsb200_batch_fetch_dense_dev(ctx, sb, ctx->kmer_cache[4].as<int32_t>(), ctx->kmer_cache[7].as<uint32_t>(), sb->cig_words + 4, &tw,
			ctx->kmer_cache[6].as<uint32_t>(), ctx->kmer_cache[5].as<int32_t>());
		if(rc == 0){
			CKK(ctx->kmer_cache[3].reserve(m * 4 + 16));
			cudaError_t e2 = cudaMemcpyAsync(ctx->kmer_cache[3].p, fall.data(), m * 4, cudaMemcpyHostToDevice, st);
			if(e2 == cudaSuccess){
				kmer_merge_kernel<<<(unsigned)((m * 32 + 255) / 256), 256, 0, st>>>(ctx->kmer_cache[3].as<uint32_t>(), (uint32_t)m, ctx->kmer_cache[4].as<int32_t>(),
					ctx->kmer_cache[5].as<int32_t>(), ctx->kmer_cache[6].as<uint32_t>(), ctx->kmer_cache[7].as<uint32_t>(), sb->d_prefix.as<uint64_t>(),
					a.results, a.status, a.ncigar, a.dense, a.dense_off, a.dense_total);
				e2 = cudaGetLastError();
				fb_launches++;
			}
			if(e2 == cudaSuccess) e2 = cudaStreamSynchronize(st);
			if(e2 != cudaSuccess){ fail(ctx, "k-mer edit: merge of the fallback pairs", e2); rc = -1; }
		}
		const std::string keep = ctx->err;
		bsb200_batch_free(ctx, sb);
		if(rc != 0){ ctx->err = keep; return done(rc); }
	}
	lap("fallback pairs");
	float h2d_ms = 0, k_ms = 0;
	CKK(cudaStreamSynchronize(st));
	cudaEventElapsedTime(&h2d_ms, ctx->ev[0], ctx->ev[1]);
	cudaEventElapsedTime(&k_ms, evk[0], evk[1]);
	b->ran = true;
	ctx->timing = bsb200_timing_t();
	uint64_t base = 0, tw = 0;
	if(turn){ base = turn->begin(); turn->taken = true; }
	int rc = fetch_impl(ctx, b, results, (cigars && !cgoff) ? cigars + base : cigars, cgoff, dense_cap > base ? dense_cap - base : 0, &tw, ncigar, status);
	if(turn) turn->end(rc == 0 ? tw : 0);
	if(total_words) *total_words = tw;
	lap("fetch");
	ctx->timing.h2d_ms = h2d_ms; ctx->timing.h2d_bytes = (segs ? segs->len0 + segs->len1 : seq_end) + n * 24;
	ctx->timing.forward_ms = k_ms; ctx->timing.forward_launches = launches;
	ctx->timing.traceback_ms = fb_ms; ctx->timing.other_launches = fb_launches;   // the fallback sub-batch (edit kernel) and the merge
	ctx->timing.run_ms = k_ms + fb_ms; ctx->timing.cells = cells; ctx->timing.waves = (uint32_t)fall.size();   // waves: pairs that took the fallback
	ctx->timing.total_ms = h2d_ms + k_ms + fb_ms + ctx->timing.d2h_ms;
	#undef CKK
	return done(rc);
}

extern "C" int bsb200_kmer_edit_batch(bsb200_ctx *ctx, uint64_t n, const uint8_t *seqs, const uint64_t *qoff, const uint32_t *qlen, const uint64_t *toff,
		const uint32_t *tlen, uint32_t ksz, bsb200_result_t *results, uint32_t *cigars, const uint64_t *cgoff, uint32_t *ncigar, int32_t *status){
	return kmer_edit_impl(ctx, n, seqs, qoff, qlen, toff, tlen, ksz, results, cigars, cgoff, 0, nullptr, ncigar, status);
}

extern "C" int bsb200_kmer_edit_batch_dense(bsb200_ctx *ctx, uint64_t n, const uint8_t *seqs, const uint64_t *qoff, const uint32_t *qlen, const uint64_t *toff,
		const uint32_t *tlen, uint32_t ksz, bsb200_result_t *results, uint32_t *cigars, uint64_t cigar_cap_words, uint64_t *total_words, uint32_t *ncigar, int32_t *status){
	if(ctx && n >= 400000 && !getenv("BSB200_NOPIPE") && ksz && seqs && qoff && qlen && toff && tlen && results)   // see dense_pipelined
		return dense_pipelined(ctx, 1, n, seqs, nullptr, qoff, qlen, toff, tlen, 0, 0, nullptr, 0, 0, 0, 0, results, cigars, cigar_cap_words, total_words, ncigar, status, ksz > 15 ? 15 : ksz);
	return kmer_edit_impl(ctx, n, seqs, qoff, qlen, toff, tlen, ksz, results, cigars, nullptr, cigar_cap_words, total_words, ncigar, status);
}

extern "C" int bsb200_kmer_edit_pairwise(bsb200_ctx *ctx, uint32_t ksz, const uint8_t *qseq, uint32_t qlen, const uint8_t *tseq, uint32_t tlen,
		bsb200_result_t *result, uint32_t *cigar, uint32_t cigar_cap, uint32_t *ncigar, int32_t *status){
	std::vector<uint8_t> arena((size_t)qlen + tlen + 1);
	if(qlen) memcpy(arena.data(), qseq, qlen);
	if(tlen) memcpy(arena.data() + qlen, tseq, tlen);
	const uint64_t qo = 0, to = qlen, cgo[2] = {0, cigar_cap};
	return kmer_edit_impl(ctx, 1, arena.data(), &qo, &qlen, &to, &tlen, ksz, result, cigar, cigar ? cgo : nullptr, 0, nullptr, ncigar, status);
}


// ---- re-alignment of reads against the MSA profile: the DP + walk of remsa_pedit_rd_bspoacore (bspoa.h:3916-4045) for batches ----------
// hdr: per job 8 ints {mlen, bw, mbeg, mend, rend, 0, 0, 0}; `in` holds per job the ten arrays of the reference's own layout (see
// remsa_kernels.cuh) at in_off[job]; match receives rend ints per job at match_off[job] (the MSA column of every read position, -1 =
// not matched), out 4 ints per job {score of the walk, status, matched positions, 0}.  matrices (may be NULL): the two DP matrices of
// every job, (2 * mlen + 1) * (bw + 2) bytes each, at mat_off[job] - for checks against the reference's own matrices.
extern "C" int bsb200_remsa_batch(bsb200_ctx *ctx, uint32_t njobs, const int32_t *hdr, const uint8_t *in, const uint64_t *in_off, uint64_t in_bytes,
		int32_t *match, const uint64_t *match_off, uint64_t match_ints, int32_t *out, uint8_t *matrices, const uint64_t *mat_off_out){
	if(!ctx || (njobs && (!hdr || !in || !in_off || !match || !match_off || !out))) return fail(ctx, "bsb200_remsa_batch", cudaSuccess);
	ctx->err.clear();
	if(njobs == 0) return 0;
	cudaSetDevice(ctx->device);
	cudaStream_t st = ctx->stream;
	const bool full = matrices && mat_off_out;
	std::vector<uint64_t> moff(njobs + 1, 0), coff(njobs + 1, 0);
	for(uint32_t j=0;j<njobs;j++){
		const int32_t *h = hdr + (size_t)j * 8;
		if(h[0] <= 0 || h[1] <= 0 || h[4] < 0) return fail(ctx, "bsb200_remsa_batch: bad job header", cudaSuccess);
		moff[j + 1] = moff[j] + (full ? ((2ull * ((uint64_t)(2 * h[0] + 1) * (uint64_t)(h[1] + 2)) + 127) / 128 * 128) : 0);
		coff[j + 1] = coff[j] + (((uint64_t)(2 * h[0] + 1) * (uint64_t)((h[1] + 31) / 32) + 7) / 8 * 8);   // 16-byte records
	}
	DevBuf *c = ctx->remsa_cache;
	CK(c[0].reserve(in_bytes + 16)); CK(c[1].reserve((size_t)njobs * 32)); CK(c[2].reserve((size_t)njobs * 8 + 8)); CK(c[3].reserve(moff[njobs] + coff[njobs] * 16 + 256, true));
	CK(c[5].reserve(match_ints * 8 + 32)); CK(c[6].reserve((size_t)njobs * 8 + 8)); CK(c[7].reserve((size_t)njobs * 16));
	cudaEventRecord(ctx->ev[0], st);
	CK(h2d_copy(ctx, c[0].p, in, in_bytes, st));
	CK(cudaMemcpyAsync(c[1].p, hdr, (size_t)njobs * 32, cudaMemcpyHostToDevice, st));
	CK(cudaMemcpyAsync(c[2].p, in_off, (size_t)njobs * 8, cudaMemcpyHostToDevice, st));
	CK(c[4].reserve((size_t)njobs * 16 + 16));
	CK(cudaMemcpyAsync(c[4].p, moff.data(), (size_t)njobs * 8, cudaMemcpyHostToDevice, st));
	CK(cudaMemcpyAsync(c[4].as<uint64_t>() + njobs, coff.data(), (size_t)njobs * 8, cudaMemcpyHostToDevice, st));
	CK(cudaMemcpyAsync(c[6].p, match_off, (size_t)njobs * 8, cudaMemcpyHostToDevice, st));
	cudaEventRecord(ctx->ev[1], st);
	RemsaArgs a;
	a.njobs = njobs; a.hdr = c[1].as<int32_t>(); a.in = c[0].as<uint8_t>(); a.in_off = c[2].as<uint64_t>();
	a.mat = c[3].as<uint8_t>(); a.mat_off = c[4].as<uint64_t>(); a.codes = (uint4*)(c[3].as<uint8_t>() + (moff[njobs] + 127) / 128 * 128); a.code_off = c[4].as<uint64_t>() + njobs;
	a.match = c[5].as<int32_t>(); a.xs = c[5].as<int32_t>() + match_ints + 4; a.match_off = c[6].as<uint64_t>(); a.out = c[7].as<int32_t>();
	if(full) CK(cudaMemsetAsync(c[3].p, 0, moff[njobs], st));   // rows outside a job's diagonals are never written (the reference's hold older calls' rows there)
	if(full) remsa_kernel<true><<<(njobs + kRemsaWarps - 1) / kRemsaWarps, kRemsaWarps * 32, 0, st>>>(a);
	else remsa_kernel<false><<<(njobs + kRemsaWarps - 1) / kRemsaWarps, kRemsaWarps * 32, 0, st>>>(a);
	CK(cudaGetLastError());
	cudaEventRecord(ctx->ev[2], st);
	CK(d2h_copy(ctx, match, c[5].p, match_ints * 4, st));
	CK(cudaMemcpyAsync(out, c[7].p, (size_t)njobs * 16, cudaMemcpyDeviceToHost, st));
	if(matrices && mat_off_out){
		for(uint32_t j=0;j<njobs;j++){
			const int32_t *h = hdr + (size_t)j * 8;
			CK(cudaMemcpyAsync(matrices + mat_off_out[j], c[3].as<uint8_t>() + moff[j], 2ull * (uint64_t)(2 * h[0] + 1) * (uint64_t)(h[1] + 2), cudaMemcpyDeviceToHost, st));
		}
	}
	cudaEventRecord(ctx->ev[3], st);
	CK(cudaStreamSynchronize(st));
	float m0 = 0, m1 = 0, m2 = 0;
	cudaEventElapsedTime(&m0, ctx->ev[0], ctx->ev[1]); cudaEventElapsedTime(&m1, ctx->ev[1], ctx->ev[2]); cudaEventElapsedTime(&m2, ctx->ev[2], ctx->ev[3]);
	ctx->timing = bsb200_timing_t();
	ctx->timing.h2d_ms = m0; ctx->timing.forward_ms = m1; ctx->timing.run_ms = m1; ctx->timing.d2h_ms = m2; ctx->timing.total_ms = m0 + m1 + m2;
	ctx->timing.forward_launches = 1; ctx->timing.h2d_bytes = in_bytes + (uint64_t)njobs * 56; ctx->timing.d2h_bytes = match_ints * 4 + (uint64_t)njobs * 16;
	{ uint64_t cells = 0; for(uint32_t j=0;j<njobs;j++){ const int32_t *h = hdr + (size_t)j * 8; cells += (uint64_t)h[1] * (2ull * (uint64_t)(h[3] - h[2])); } ctx->timing.cells = cells; }
	ctx->timing.trace_bytes = moff[njobs] + coff[njobs] * 16;
	return 0;
}

// ---- host helpers of the multi-GPU split (bsalign_b200/shard.py): compact arenas per shard, pair-ordered merge of the shards' cigars ----
// The pairs idx[0..m) of a batch are copied into one compact arena (query then target of each pair, in idx order); out_qoff / out_toff
// receive their offsets in it.  Returns the bytes written (out_seqs may be NULL to size the arena).
extern "C" uint64_t bsb200_pack_pairs(const uint8_t *seqs, const uint64_t *qoff, const uint32_t *qlen, const uint64_t *toff, const uint32_t *tlen,
		const uint64_t *idx, uint64_t m, uint8_t *out_seqs, uint64_t *out_qoff, uint64_t *out_toff, int nthreads){
	std::vector<uint64_t> pos(m + 1, 0);
	for(uint64_t k=0;k<m;k++) pos[k + 1] = pos[k] + qlen[idx[k]] + tlen[idx[k]];
	if(out_qoff) for(uint64_t k=0;k<m;k++) out_qoff[k] = pos[k];
	if(out_toff) for(uint64_t k=0;k<m;k++) out_toff[k] = pos[k] + qlen[idx[k]];
	if(!out_seqs) return pos[m];
	if(nthreads < 1) nthreads = 1;
	auto work = [&](int w){
		// equal byte shares
		const uint64_t lo_b = pos[m] / nthreads * w, hi_b = w + 1 == nthreads ? pos[m] : pos[m] / nthreads * (w + 1);
		uint64_t k = std::lower_bound(pos.begin(), pos.end() - 1, lo_b) - pos.begin();
		for(;k<m && pos[k]<hi_b;k++){
			const uint64_t i = idx[k];
			if(qlen[i]) memcpy(out_seqs + pos[k], seqs + qoff[i], qlen[i]);
			if(tlen[i]) memcpy(out_seqs + pos[k] + qlen[i], seqs + toff[i], tlen[i]);
		}
	};
	std::vector<std::thread> th;
	for(int w=1;w<nthreads;w++) th.emplace_back(work, w);
	work(0);
	for(auto &t : th) t.join();
	return pos[m];
}

// The same packing ON THE DEVICE: the caller's arena is already in this device's memory as it is (one large copy instead of a host-side
// gather); one warp per sequence moves it to its place in the compact arena (query then target of each pair, in idx order).
__global__ void __launch_bounds__(256) pack_pairs_kernel(const uint8_t *src, uint8_t *dst, const uint64_t *soff, const uint64_t *doff, const uint32_t *len, uint64_t nseg){
	const uint64_t sgm = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
	if(sgm >= nseg) return;
	const uint32_t lane = threadIdx.x & 31, n = len[sgm];
	const uint8_t *s_ = src + soff[sgm]; uint8_t *d_ = dst + doff[sgm];
	// bytes up to the destination's 16-byte line, then 16 bytes per lane and step when the source is aligned the same way
	const uint32_t head = (uint32_t)((16 - ((uintptr_t)d_ & 15)) & 15);
	if((((uintptr_t)s_ ^ (uintptr_t)d_) & 15) == 0 && n >= 64){
		const uint32_t h = head < n ? head : n;
		if(lane < h) d_[lane] = s_[lane];
		const uint32_t body = (n - h) / 16;
		const uint4 *s4 = (const uint4*)(s_ + h); uint4 *d4 = (uint4*)(d_ + h);
		for(uint32_t k=lane;k<body;k+=32) d4[k] = s4[k];
		for(uint32_t k=h+body*16+lane;k<n;k+=32) d_[k] = s_[k];
	} else for(uint32_t k=lane;k<n;k+=32) d_[k] = s_[k];
}

extern "C" int bsb200_pack_pairs_dev(bsb200_ctx *ctx, const uint8_t *d_src, const uint64_t *qoff, const uint32_t *qlen, const uint64_t *toff, const uint32_t *tlen,
		const uint64_t *idx, uint64_t m, uint8_t *d_dst){
	if(!ctx || (m && (!d_src || !d_dst || !qoff || !qlen || !toff || !tlen || !idx))) return fail(ctx, "bsb200_pack_pairs_dev", cudaSuccess);
	ctx->err.clear();
	if(m == 0) return 0;
	cudaSetDevice(ctx->device);
	HostBuf &hb = ctx->host_cache[6];
	DevBuf &db = ctx->kmer_cache[3];
	const uint64_t nseg = 2 * m;
	CK(hb.reserve(nseg * 20)); CK(db.reserve(nseg * 20));
	uint64_t *so = hb.as<uint64_t>(), *dof = so + nseg; uint32_t *ln = (uint32_t*)(dof + nseg);
	uint64_t pos = 0;
	for(uint64_t k=0;k<m;k++){
		const uint64_t i = idx[k];
		so[2 * k] = qoff[i]; dof[2 * k] = pos; ln[2 * k] = qlen[i]; pos += qlen[i];
		so[2 * k + 1] = toff[i]; dof[2 * k + 1] = pos; ln[2 * k + 1] = tlen[i]; pos += tlen[i];
	}
	CK(cudaMemcpyAsync(db.p, hb.p, nseg * 20, cudaMemcpyHostToDevice, ctx->stream));
	const uint64_t *dso = db.as<uint64_t>(), *ddo = dso + nseg; const uint32_t *dln = (const uint32_t*)(ddo + nseg);
	pack_pairs_kernel<<<(unsigned)((nseg * 32 + 255) / 256), 256, 0, ctx->stream>>>(d_src, d_dst, dso, ddo, dln, nseg);
	CK(cudaGetLastError());
	CK(cudaStreamSynchronize(ctx->stream));
	return 0;
}

// dst[dst_off[k] .. + len[k]) = src[src_off[k] .. + len[k]) for k in [0, m): 32-bit words
extern "C" void bsb200_scatter_words(uint32_t *dst, const uint64_t *dst_off, const uint32_t *src, const uint64_t *src_off, const uint32_t *len, uint64_t m, int nthreads){
	if(nthreads < 1) nthreads = 1;
	auto work = [&](int w){
		for(uint64_t k=m*w/nthreads;k<m*(w+1)/nthreads;k++) if(len[k]) memcpy(dst + dst_off[k], src + src_off[k], (size_t)len[k] * 4);
	};
	std::vector<std::thread> th;
	for(int w=1;w<nthreads;w++) th.emplace_back(work, w);
	work(0);
	for(auto &t : th) t.join();
}

extern "C" void bsb200_batch_free(bsb200_ctx *ctx, bsb200_batch *b){
	if(!b) return;
	if(ctx){ cudaSetDevice(ctx->device); cudaStreamSynchronize(ctx->stream); }
	DevBuf *ds[] = {&b->d_seqs, &b->d_qoff, &b->d_toff, &b->d_qlen, &b->d_tlen, &b->d_order, &b->d_trace_off, &b->d_results, &b->d_status,
		&b->d_ncigar, &b->d_cig_raw, &b->d_cig_off, &b->d_cig_dense, &b->d_dense_off, &b->d_dense_total, &b->d_block_rows, &b->d_prefix, &b->d_bits};
	HostBuf *hs[] = {&b->h_results, &b->h_status, &b->h_ncigar, &b->h_dense_off, &b->h_dense, &b->h_total};
	if(ctx){ // park the allocations for the next batch (a cache slot that is still occupied keeps the larger buffer)
		for(int k=0;k<18;k++){ if(ctx->dev_cache[k].cap < ds[k]->cap){ ctx->dev_cache[k].release(); ctx->dev_cache[k] = *ds[k]; } else ds[k]->release(); }
		for(int k=0;k<6;k++){ if(ctx->host_cache[k].cap < hs[k]->cap){ ctx->host_cache[k].release(); ctx->host_cache[k] = *hs[k]; } else hs[k]->release(); }
	} else {
		for(auto d : ds) d->release();
		for(auto h : hs) h->release();
	}
	delete b;
}

static int run_whole(bsb200_ctx *ctx, int kind, uint64_t n, const uint8_t *seqs,
		const uint64_t *qoff, const uint32_t *qlen, const uint64_t *toff, const uint32_t *tlen,
		int mode, uint32_t bandwidth, const int8_t *matrix, int8_t go1, int8_t ge1, int8_t go2, int8_t ge2,
		bsb200_result_t *results, uint32_t *cigars, const uint64_t *cgoff, uint32_t *ncigar, int32_t *status){
	if(!ctx) return -1;
	bsb200_batch *b = bsb200_batch_upload(ctx, kind, n, seqs, qoff, qlen, toff, tlen, mode, bandwidth, matrix, go1, ge1, go2, ge2, cigars && cgoff);
	if(!b) return -1;
	int rc = bsb200_batch_run(ctx, b);
	if(rc == 0) rc = bsb200_batch_fetch(ctx, b, results, cigars, cgoff, ncigar, status);
	bsb200_timing_t tm = ctx->timing;
	bsb200_batch_free(ctx, b);
	tm.total_ms = tm.h2d_ms + tm.run_ms + tm.d2h_ms;
	ctx->timing = tm;
	return rc;
}

extern "C" int bsb200_epi8_pairwise_batch(bsb200_ctx *ctx, uint64_t n, const uint8_t *seqs,
		const uint64_t *qoff, const uint32_t *qlen, const uint64_t *toff, const uint32_t *tlen,
		int mode, uint32_t bandwidth, const int8_t matrix[16], int8_t go1, int8_t ge1, int8_t go2, int8_t ge2,
		bsb200_result_t *results, uint32_t *cigars, const uint64_t *cgoff, uint32_t *ncigar, int32_t *status){
	return run_whole(ctx, 0, n, seqs, qoff, qlen, toff, tlen, mode, bandwidth, matrix, go1, ge1, go2, ge2, results, cigars, cgoff, ncigar, status);
}

extern "C" int bsb200_edit_pairwise_batch(bsb200_ctx *ctx, uint64_t n, const uint8_t *seqs,
		const uint64_t *qoff, const uint32_t *qlen, const uint64_t *toff, const uint32_t *tlen,
		int mode, uint32_t bandwidth,
		bsb200_result_t *results, uint32_t *cigars, const uint64_t *cgoff, uint32_t *ncigar, int32_t *status){
	return run_whole(ctx, 1, n, seqs, qoff, qlen, toff, tlen, mode, bandwidth, nullptr, 0, 0, 0, 0, results, cigars, cgoff, ncigar, status);
}

// ---- one call, dense pair-ordered cigars (the fast path of the staged form as ONE ABI function) ---------------------------------
// Large EDIT batches are cut into chunks of consecutive pairs that two host threads push through two contexts of the same device
// (ctx and ctx->helper, one stream each): the H2D copy of chunk c + 1 and the host plan of chunk c + 2 run while chunk c is in its
// kernel and its results travel back.  A million 300 bp pairs are 600 MB in, 5 ms of kernel and 100 MB out: without the overlap the
// call takes the sum, with it little more than the copy.  The chunks' dense cigars are fetched in chunk order, each behind the
// words of its predecessors.  Either the byte arena (seqs) or the 2-bit words (bits; offsets are base offsets) is given.
static int dense_pipelined(bsb200_ctx *ctx, int kind, uint64_t n, const uint8_t *seqs, const uint64_t *bits,
		const uint64_t *qoff, const uint32_t *qlen, const uint64_t *toff, const uint32_t *tlen,
		int mode, uint32_t bandwidth, const int8_t *matrix, int8_t go1, int8_t ge1, int8_t go2, int8_t ge2,
		bsb200_result_t *results, uint32_t *cigars, uint64_t cigar_cap_words, uint64_t *total_words, uint32_t *ncigar, int32_t *status, uint32_t ksz){
	if(!ctx->helper){
		ctx->helper = bsb200_create(ctx->device, ctx->trace_budget ? ctx->trace_budget / 2 : 0);
		if(!ctx->helper) return fail(ctx, "second context for the pipelined call", cudaSuccess);
	}
	ctx->err.clear();
	const auto call0 = std::chrono::steady_clock::now();
	// (the k-mer guided edit, ksz > 0, is kernel bound and its kernel likes many pairs per launch: fewer, larger chunks)
	const uint64_t K = ksz ? std::min<uint64_t>(4, std::max<uint64_t>(2, n / 200000)) : std::min<uint64_t>(8, std::max<uint64_t>(2, n / 131072));
	struct { std::mutex m, upload; std::condition_variable cv; uint64_t next_fetch = 0, base = 0; int err = 0; std::string msg; } sy;
	bsb200_timing_t acc[2] = {bsb200_timing_t(), bsb200_timing_t()};
	// pinned or pageable caller arrays: asked once here, not once per chunk and copy (the query costs ~30 us)
	const int src_pg = is_pageable(seqs ? (const void*)seqs : (const void*)bits) ? 1 : 0;
	const int dst_pg = (is_pageable(results) || (cigars && is_pageable(cigars))) ? 1 : 0;
	auto worker = [&](int t){
		bsb200_ctx *cx = t ? ctx->helper : ctx;
		cudaSetDevice(cx->device);
		tl_src_pageable = src_pg; tl_dst_pageable = dst_pg;
		struct Reset { ~Reset(){ tl_src_pageable = -1; tl_dst_pageable = -1; } } reset_;
		// rebased offsets live in pinned memory: a copy from pageable memory would wait for the sequence copy queued before it
		HostBuf &hq2 = cx->host_cache[6], &ht2 = cx->host_cache[7];
		// the stretch of the arena a chunk's queries come from and the one its targets come from: one copy when they touch or overlap
		// (pairs stored q, t, q, t ...), two when the arena holds all queries first and all targets behind them.  The chunk after next is
		// prepared while this worker's chunk is still on the device.
		SeqSegs sg; bool two = false; uint64_t lo = 0;
		auto prep = [&](uint64_t c){
			const uint64_t c0 = n * c / K, c1 = n * (c + 1) / K, m = c1 - c0;
			uint64_t qlo = ~0ull, qhi = 0, tlo = ~0ull, thi = 0;
			for(uint64_t i=c0;i<c1;i++){
				qlo = std::min(qlo, qoff[i]); qhi = std::max(qhi, qoff[i] + qlen[i]);
				tlo = std::min(tlo, toff[i]); thi = std::max(thi, toff[i] + tlen[i]);
			}
			if(bits){ qlo &= ~31ull; tlo &= ~31ull; }   // whole words
			if(hq2.reserve(m * 8) != cudaSuccess || ht2.reserve(m * 8) != cudaSuccess){ cudaGetLastError(); std::lock_guard<std::mutex> g(sy.m); if(!sy.err){ sy.err = -1; sy.msg = "pinned host allocation"; } return; }
			uint64_t *q2 = hq2.as<uint64_t>(), *t2 = ht2.as<uint64_t>();
			lo = std::min(qlo, tlo);
			const uint64_t gap = qlo < tlo ? (tlo > qhi ? tlo - qhi : 0) : (qlo > thi ? qlo - thi : 0);
			two = gap > 4096;
			if(two){
				const bool qfirst = qlo < tlo;
				const uint64_t len0 = qfirst ? qhi - qlo : thi - tlo, at1 = (len0 + 127) / 128 * 128;
				sg.len0 = len0; sg.len1 = qfirst ? thi - tlo : qhi - qlo; sg.src1 = (qfirst ? tlo : qlo) - lo; sg.at1 = at1;
				for(uint64_t i=0;i<m;i++){
					q2[i] = qfirst ? qoff[c0 + i] - qlo : qoff[c0 + i] - qlo + at1;
					t2[i] = qfirst ? toff[c0 + i] - tlo + at1 : toff[c0 + i] - tlo;
				}
			} else for(uint64_t i=0;i<m;i++){ q2[i] = qoff[c0 + i] - lo; t2[i] = toff[c0 + i] - lo; }
		};
		if((uint64_t)t < K) prep((uint64_t)t);
		for(uint64_t c=(uint64_t)t;c<K;c+=2){
			const uint64_t c0 = n * c / K, c1 = n * (c + 1) / K, m = c1 - c0;
			const auto tstart = std::chrono::steady_clock::now();
			uint64_t *q2 = hq2.as<uint64_t>(), *t2 = ht2.as<uint64_t>();
			const SeqSegs sgc = sg; const SeqSegs *psg = two ? &sgc : nullptr;
			int rc = 0;
			{ std::lock_guard<std::mutex> g(sy.m); rc = sy.err; }
			auto w0 = tstart;
			auto wlap = [&](const char *what){ if(getenv("BSB200_HOSTPROF")){ auto w1 = std::chrono::steady_clock::now(); fprintf(stderr, "[pipe %d] chunk %llu %-8s %7.3f ms  (at %7.3f)\n", t, (unsigned long long)c, what, std::chrono::duration<double, std::milli>(w1 - w0).count(), std::chrono::duration<double, std::milli>(w1 - call0).count()); w0 = w1; } };
			if(ksz){   // the k-mer guided edit is one function: copies, kernel rounds, fallback pairs; it takes its fetch turn through the hook
				FetchTurn turn;
				turn.begin = [&]() -> uint64_t { std::unique_lock<std::mutex> lk(sy.m); sy.cv.wait(lk, [&]{ return sy.next_fetch == c; }); return sy.base; };
				turn.end = [&](uint64_t tw){ { std::lock_guard<std::mutex> g(sy.m); sy.base += tw; sy.next_fetch = c + 1; } sy.cv.notify_all(); };
				if(rc == 0) rc = kmer_edit_impl(cx, m, seqs + lo, q2, qlen + c0, t2, tlen + c0, ksz, results + c0, cigars, nullptr, cigar_cap_words, nullptr,
					ncigar ? ncigar + c0 : nullptr, status ? status + c0 : nullptr, psg, &turn, 2);
				if(!turn.taken){ turn.begin(); turn.end(0); }
				if(rc){ std::lock_guard<std::mutex> g(sy.m); if(!sy.err){ sy.err = rc; sy.msg = cx->err; } }
				const bsb200_timing_t tm = cx->timing;
				bsb200_timing_t &A = acc[t];
				A.h2d_ms += tm.h2d_ms; A.forward_ms += tm.forward_ms; A.traceback_ms += tm.traceback_ms; A.d2h_ms += tm.d2h_ms; A.run_ms += tm.run_ms;
				A.forward_launches += tm.forward_launches; A.other_launches += tm.other_launches; A.cells += tm.cells; A.h2d_bytes += tm.h2d_bytes; A.d2h_bytes += tm.d2h_bytes;
				A.reserved += tm.waves;   // pairs that took the fallback
				if(c + 2 < K) prep(c + 2);
				continue;
			}
			// one upload at a time: the link is shared anyway, and two workers that copy side by side fall into lock step (both copy, both
			// compute, both fetch) instead of one copying while the other computes
			bsb200_batch *b = nullptr;
			{
				std::lock_guard<std::mutex> up(sy.upload);
				wlap("token");
				b = rc ? nullptr : upload_impl(cx, kind, m, seqs ? seqs + lo : nullptr, nullptr, bits ? bits + (lo >> 5) : nullptr, q2, qlen + c0, t2, tlen + c0,
					mode, bandwidth, matrix, go1, ge1, go2, ge2, cigars != nullptr, psg);
			}
			if(!b && !rc) rc = -1;
			bsb200_timing_t tm = cx->timing;
			wlap("upload");
			if(rc == 0){ rc = bsb200_batch_run(cx, b); const float h = tm.h2d_ms; const uint64_t hb = tm.h2d_bytes; tm = cx->timing; tm.h2d_ms = h; tm.h2d_bytes = hb; }
			wlap("run");
			if(c + 2 < K) prep(c + 2);
			wlap("prep+2");
			std::unique_lock<std::mutex> lk(sy.m);
			sy.cv.wait(lk, [&]{ return sy.next_fetch == c; });   // the dense cigars of the chunks go out in chunk order
			wlap("wait");
			if(rc == 0 && sy.err) rc = sy.err;
			const uint64_t base = sy.base;
			lk.unlock();
			uint64_t tw = 0;
			if(rc == 0){
				rc = fetch_impl(cx, b, results + c0, cigars ? cigars + base : nullptr, nullptr, cigar_cap_words > base ? cigar_cap_words - base : 0, &tw, ncigar ? ncigar + c0 : nullptr, status ? status + c0 : nullptr);
				tm.d2h_ms = cx->timing.d2h_ms; tm.d2h_bytes = cx->timing.d2h_bytes;
			}
			wlap("fetch");
			lk.lock();
			if(rc && !sy.err){ sy.err = rc; sy.msg = cx->err; }
			sy.base = base + tw; sy.next_fetch = c + 1;
			lk.unlock();
			sy.cv.notify_all();
			if(b) bsb200_batch_free(cx, b);
			wlap("free");
			bsb200_timing_t &A = acc[t];
			A.h2d_ms += tm.h2d_ms; A.forward_ms += tm.forward_ms; A.traceback_ms += tm.traceback_ms; A.d2h_ms += tm.d2h_ms; A.run_ms += tm.run_ms;
			A.forward_launches += tm.forward_launches; A.traceback_launches += tm.traceback_launches; A.other_launches += tm.other_launches;
			A.cells += tm.cells; A.trace_bytes += tm.trace_bytes; A.h2d_bytes += tm.h2d_bytes; A.d2h_bytes += tm.d2h_bytes;
		}
	};
	std::thread th(worker, 1);
	worker(0);
	th.join();
	if(getenv("BSB200_HOSTPROF")) fprintf(stderr, "[pipe] all chunks done at %7.3f ms\n", std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - call0).count());
	cudaSetDevice(ctx->device);
	bsb200_timing_t T = acc[0];
	T.h2d_ms += acc[1].h2d_ms; T.forward_ms += acc[1].forward_ms; T.traceback_ms += acc[1].traceback_ms; T.d2h_ms += acc[1].d2h_ms; T.run_ms += acc[1].run_ms;
	T.forward_launches += acc[1].forward_launches; T.traceback_launches += acc[1].traceback_launches; T.other_launches += acc[1].other_launches;
	T.cells += acc[1].cells; T.trace_bytes += acc[1].trace_bytes; T.h2d_bytes += acc[1].h2d_bytes; T.d2h_bytes += acc[1].d2h_bytes;
	T.waves = ksz ? acc[0].reserved + acc[1].reserved : (uint32_t)K; T.reserved = 0;
	T.total_ms = T.h2d_ms + T.run_ms + T.d2h_ms;   // sums over the chunks: they overlap, the wall time is shorter
	ctx->timing = T;
	if(total_words) *total_words = sy.base;
	if(sy.err){ ctx->err = sy.msg; return sy.err; }
	return 0;
}

// Large edit batches of SHORT pairs: the two contexts each hold the traceback store of a chunk, so the whole batch's store must be a small
// part of the device (long pairs go through the plain path, whose waves are planned against all of the free memory).
static bool pipeline_pays(const bsb200_ctx *ctx, int kind, uint64_t n, const uint32_t *qlen, const uint32_t *tlen, int mode, uint32_t bandwidth){
	if(getenv("BSB200_NOPIPE") || kind != 1 || n < 262144 || !qlen || !tlen) return false;
	const int NT = plan_threads(n);
	std::vector<uint64_t> part(NT, 0);
	par_slices(n, NT, [&](int w, uint64_t lo, uint64_t hi){
		uint64_t b = 0;
		for(uint64_t i=lo;i<hi;i++) if(qlen[i] && tlen[i]) b += edit_trace_bytes(edit_bandwidth(qlen[i], tlen[i], mode & 3, bandwidth), tlen[i]);
		part[w] = b;
	});
	uint64_t total = 0;
	for(uint64_t b : part) total += b;
	return total <= ctx->total_mem / 4;
}

extern "C" int bsb200_pairwise_batch_dense(bsb200_ctx *ctx, int kind, uint64_t n, const uint8_t *seqs,
		const uint64_t *qoff, const uint32_t *qlen, const uint64_t *toff, const uint32_t *tlen,
		int mode, uint32_t bandwidth, const int8_t matrix[16], int8_t go1, int8_t ge1, int8_t go2, int8_t ge2,
		bsb200_result_t *results, uint32_t *cigars, uint64_t cigar_cap_words, uint64_t *total_words, uint32_t *ncigar, int32_t *status){
	if(!ctx) return -1;
	if(seqs && qoff && toff && results && pipeline_pays(ctx, kind, n, qlen, tlen, mode, bandwidth))
		return dense_pipelined(ctx, kind, n, seqs, nullptr, qoff, qlen, toff, tlen, mode, bandwidth, matrix, go1, ge1, go2, ge2, results, cigars, cigar_cap_words, total_words, ncigar, status, 0);
	bsb200_batch *b = bsb200_batch_upload(ctx, kind, n, seqs, qoff, qlen, toff, tlen, mode, bandwidth, matrix, go1, ge1, go2, ge2, cigars != nullptr);
	if(!b) return -1;
	int rc = bsb200_batch_run(ctx, b);
	if(rc == 0) rc = bsb200_batch_fetch_dense(ctx, b, results, cigars, cigar_cap_words, total_words, ncigar, status);
	bsb200_timing_t tm = ctx->timing;
	bsb200_batch_free(ctx, b);
	tm.total_ms = tm.h2d_ms + tm.run_ms + tm.d2h_ms;
	ctx->timing = tm;
	return rc;
}

// the same call with the sequences 2-bit packed (a BaseBank's words, see bsb200_batch_upload_bits; qoff / toff are base offsets)
extern "C" int bsb200_pairwise_batch_dense_bits(bsb200_ctx *ctx, int kind, uint64_t n, const uint64_t *bits,
		const uint64_t *qoff, const uint32_t *qlen, const uint64_t *toff, const uint32_t *tlen,
		int mode, uint32_t bandwidth, const int8_t matrix[16], int8_t go1, int8_t ge1, int8_t go2, int8_t ge2,
		bsb200_result_t *results, uint32_t *cigars, uint64_t cigar_cap_words, uint64_t *total_words, uint32_t *ncigar, int32_t *status){
	if(!ctx) return -1;
	if(bits && qoff && toff && results && pipeline_pays(ctx, kind, n, qlen, tlen, mode, bandwidth))
		return dense_pipelined(ctx, kind, n, nullptr, bits, qoff, qlen, toff, tlen, mode, bandwidth, matrix, go1, ge1, go2, ge2, results, cigars, cigar_cap_words, total_words, ncigar, status, 0);
	bsb200_batch *b = bsb200_batch_upload_bits(ctx, kind, n, bits, qoff, qlen, toff, tlen, mode, bandwidth, matrix, go1, ge1, go2, ge2, cigars != nullptr);
	if(!b) return -1;
	int rc = bsb200_batch_run(ctx, b);
	if(rc == 0) rc = bsb200_batch_fetch_dense(ctx, b, results, cigars, cigar_cap_words, total_words, ncigar, status);
	bsb200_timing_t tm = ctx->timing;
	bsb200_batch_free(ctx, b);
	tm.total_ms = tm.h2d_ms + tm.run_ms + tm.d2h_ms;
	ctx->timing = tm;
	return rc;
}

// ---- one pointer per sequence, as the reference hands them out (bsalign.h:399: u1i *qseq, u1i *tseq per call) -----------------------
// The pairs are gathered into one pinned arena by `nthreads` host threads and go through the arena entry point.  cigar_out[i] (may be
// NULL) must have room for qlen[i] + tlen[i] + 2 words.
extern "C" int bsb200_pairwise_batch_ptrs(bsb200_ctx *ctx, int kind, uint64_t n, const uint8_t *const *q, const uint32_t *qlen,
		const uint8_t *const *t, const uint32_t *tlen, int mode, uint32_t bandwidth, const int8_t matrix[16],
		int8_t go1, int8_t ge1, int8_t go2, int8_t ge2, bsb200_result_t *results, uint32_t *const *cigar_out, uint32_t *ncigar, int32_t *status, int nthreads){
	if(!ctx) return -1;
	if(n && (!q || !t || !qlen || !tlen)) return fail(ctx, "bsb200_pairwise_batch_ptrs", cudaSuccess);
	cudaSetDevice(ctx->device);
	std::vector<uint64_t> qoff(n), toff(n), cgoff(n + 1, 0);
	uint64_t pos = 0;
	for(uint64_t i=0;i<n;i++){ qoff[i] = pos; pos += qlen[i]; toff[i] = pos; pos += tlen[i]; cgoff[i + 1] = cgoff[i] + ((qlen[i] && tlen[i]) ? (uint64_t)qlen[i] + tlen[i] + 2 : 0); }
	HostBuf &arena = ctx->host_cache[6];
	CK(arena.reserve(pos + 16));
	uint8_t *dst = arena.as<uint8_t>();
	if(nthreads < 1) nthreads = 1;
	auto work = [&](int w){
		for(uint64_t i=n*w/nthreads;i<n*(w+1)/nthreads;i++){
			if(qlen[i]) memcpy(dst + qoff[i], q[i], qlen[i]);
			if(tlen[i]) memcpy(dst + toff[i], t[i], tlen[i]);
		}
	};
	{ std::vector<std::thread> th; for(int w=1;w<nthreads;w++) th.emplace_back(work, w); work(0); for(auto &x : th) x.join(); }
	const bool want = cigar_out != nullptr;
	std::vector<uint32_t> cg(want ? cgoff[n] : 0), ncg(n);
	int rc = run_whole(ctx, kind, n, dst, qoff.data(), qlen, toff.data(), tlen, mode, bandwidth, matrix, go1, ge1, go2, ge2,
		results, want ? cg.data() : nullptr, want ? cgoff.data() : nullptr, ncg.data(), status);
	if(rc) return rc;
	if(ncigar) memcpy(ncigar, ncg.data(), n * 4);
	if(want) for(uint64_t i=0;i<n;i++) if(cigar_out[i] && ncg[i]) memcpy(cigar_out[i], cg.data() + cgoff[i], (size_t)ncg[i] * 4);
	return 0;
}

// ---- every GPU of the box from ONE process (the reference is a single-process C library): one context per device, the batch cut
// into shards of equal DP cells (heaviest first, dealt in snake order), one host thread per device packs its shard's compact arena
// and runs it through the arena entry point; results land in the caller's arrays at the pairs' own indices ------------------------
extern "C" int bsb200_pairwise_batch_multi(bsb200_ctx *const *ctxs, int nctx, int kind, uint64_t n, const uint8_t *seqs,
		const uint64_t *qoff, const uint32_t *qlen, const uint64_t *toff, const uint32_t *tlen,
		int mode, uint32_t bandwidth, const int8_t matrix[16], int8_t go1, int8_t ge1, int8_t go2, int8_t ge2,
		bsb200_result_t *results, uint32_t *cigars, const uint64_t *cgoff, uint32_t *ncigar, int32_t *status){
	if(!ctxs || nctx < 1) return -1;
	for(int d=0;d<nctx;d++) if(!ctxs[d]) return -1;
	if(nctx == 1) return run_whole(ctxs[0], kind, n, seqs, qoff, qlen, toff, tlen, mode, bandwidth, matrix, go1, ge1, go2, ge2, results, cigars, cgoff, ncigar, status);
	// balanced partition (the same rule as bsalign_b200/shard.py: nominal band cells, heaviest first, snake order)
	std::vector<std::pair<uint64_t, uint64_t>> wk(n);
	for(uint64_t i=0;i<n;i++){
		const uint64_t bw = kind == 0 ? bsb200_epi8_bandwidth(qlen[i], bandwidth) : edit_bandwidth(qlen[i], tlen[i], mode, bandwidth);
		wk[i] = std::make_pair(~(bw * (uint64_t)tlen[i]), i);
	}
	std::sort(wk.begin(), wk.end());
	std::vector<std::vector<uint64_t>> part(nctx);
	for(uint64_t k=0;k<n;k++){ const uint64_t rnd = k / nctx, c = k % nctx; part[(rnd & 1) ? nctx - 1 - c : c].push_back(wk[k].second); }
	std::vector<int> rcs(nctx, 0);
	auto work = [&](int d){
		std::vector<uint64_t> &idx = part[d];
		std::sort(idx.begin(), idx.end());
		const uint64_t m = idx.size();
		std::vector<uint64_t> sq(m), st_(m), scg(m + 1, 0);
		std::vector<uint32_t> sql(m), stl(m), sncg(m);
		std::vector<int32_t> sst(m);
		std::vector<bsb200_result_t> sres(m);
		const uint64_t bytes = bsb200_pack_pairs(seqs, qoff, qlen, toff, tlen, idx.data(), m, nullptr, nullptr, nullptr, 1);
		std::vector<uint8_t> arena(bytes + 16);
		bsb200_pack_pairs(seqs, qoff, qlen, toff, tlen, idx.data(), m, arena.data(), sq.data(), st_.data(), 2);
		for(uint64_t k=0;k<m;k++){ sql[k] = qlen[idx[k]]; stl[k] = tlen[idx[k]]; scg[k + 1] = scg[k] + ((cigars && cgoff) ? cgoff[idx[k] + 1] - cgoff[idx[k]] : 0); }
		std::vector<uint32_t> scig((cigars && cgoff) ? scg[m] : 0);
		rcs[d] = run_whole(ctxs[d], kind, m, arena.data(), sq.data(), sql.data(), st_.data(), stl.data(), mode, bandwidth, matrix, go1, ge1, go2, ge2,
			sres.data(), (cigars && cgoff) ? scig.data() : nullptr, (cigars && cgoff) ? scg.data() : nullptr, sncg.data(), sst.data());
		if(rcs[d]) return;
		for(uint64_t k=0;k<m;k++){
			const uint64_t i = idx[k];
			if(results) results[i] = sres[k];
			if(ncigar) ncigar[i] = sncg[k];
			if(status) status[i] = sst[k];
			if(cigars && cgoff){ const uint64_t c = std::min<uint64_t>(sncg[k], scg[k + 1] - scg[k]); if(c) memcpy(cigars + cgoff[i], scig.data() + scg[k], c * 4); }
		}
	};
	{ std::vector<std::thread> th; for(int d=1;d<nctx;d++) th.emplace_back(work, d); work(0); for(auto &x : th) x.join(); }
	for(int d=0;d<nctx;d++) if(rcs[d]) return rcs[d];
	return 0;
}

static int run_single(bsb200_ctx *ctx, int kind, const uint8_t *qseq, uint32_t qlen, const uint8_t *tseq, uint32_t tlen,
		int mode, uint32_t bandwidth, const int8_t *matrix, int8_t go1, int8_t ge1, int8_t go2, int8_t ge2,
		bsb200_result_t *result, uint32_t *cigar, uint32_t cigar_cap, uint32_t *ncigar, int32_t *status){
	std::vector<uint8_t> seqs((size_t)qlen + tlen + 1);
	if(qlen) memcpy(seqs.data(), qseq, qlen);
	if(tlen) memcpy(seqs.data() + qlen, tseq, tlen);
	uint64_t qoff = 0, toff = qlen, cgoff[2] = {0, cigar_cap};
	return run_whole(ctx, kind, 1, seqs.data(), &qoff, &qlen, &toff, &tlen, mode, bandwidth, matrix, go1, ge1, go2, ge2,
		result, cigar, cigar ? cgoff : nullptr, ncigar, status);
}

extern "C" int bsb200_epi8_pairwise(bsb200_ctx *ctx, const uint8_t *qseq, uint32_t qlen, const uint8_t *tseq, uint32_t tlen,
		int mode, uint32_t bandwidth, const int8_t matrix[16], int8_t go1, int8_t ge1, int8_t go2, int8_t ge2,
		bsb200_result_t *result, uint32_t *cigar, uint32_t cigar_cap, uint32_t *ncigar, int32_t *status){
	return run_single(ctx, 0, qseq, qlen, tseq, tlen, mode, bandwidth, matrix, go1, ge1, go2, ge2, result, cigar, cigar_cap, ncigar, status);
}

extern "C" int bsb200_edit_pairwise(bsb200_ctx *ctx, const uint8_t *qseq, uint32_t qlen, const uint8_t *tseq, uint32_t tlen,
		int mode, uint32_t bandwidth,
		bsb200_result_t *result, uint32_t *cigar, uint32_t cigar_cap, uint32_t *ncigar, int32_t *status){
	return run_single(ctx, 1, qseq, qlen, tseq, tlen, mode, bandwidth, nullptr, 0, 0, 0, 0, result, cigar, cigar_cap, ncigar, status);
}

// ---- development aid: copy one pair's raw traceback block (epi8) back to the host -------------------
// Layout: (tlen+1) rows of (pw+1) array images (8 regions of S bytes, see common.cuh), then (tlen+1) anchor
// records of 20 ints {ub[17], rbeg, 0, 0}.  Valid after bsb200_batch_run for single-wave batches.
extern "C" int64_t bsb200_debug_trace(bsb200_ctx *ctx, bsb200_batch *b, uint64_t pair, uint8_t *out, uint64_t cap, uint32_t *bw_out, int *pw_out, int *ubias_out){
	if(!ctx || !b || b->kind != 0 || pair >= b->n || b->waves.size() != 1 || b->empty[pair]) return -1;
	std::vector<uint32_t> ql(1), tl(1);
	cudaMemcpy(ql.data(), b->d_qlen.as<uint32_t>() + pair, 4, cudaMemcpyDeviceToHost);
	cudaMemcpy(tl.data(), b->d_tlen.as<uint32_t>() + pair, 4, cudaMemcpyDeviceToHost);
	uint32_t bw = bsb200_epi8_bandwidth(ql[0], b->bandwidth);
	uint64_t bytes = ((uint64_t)((b->wave_split ? epi8_wave_use_anchors(b->max_bw / 16) : epi8_use_anchors(b->max_bw / 16)) ? epi8_row_bytes(bw / 16, b->pw) : epi8_image_bytes(bw / 16) * (b->pw + 1)) + kMetaInts * 4) * ((uint64_t)tl[0] + 1 + (b->wave_split ? epi8_wave_slack((uint32_t)b->wave_split) : 0u));   // (skewed layout when the wavefront kernel wrote it: epi8_wave.cuh)
	if(bytes > cap) return -(int64_t)bytes;
	if(cudaMemcpy(out, ctx->trace.as<uint8_t>() + b->trace_off[pair], bytes, cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
	if(bw_out) *bw_out = bw;
	if(pw_out) *pw_out = b->pw;
	if(ubias_out) *ubias_out = (b->pw >= 1 && b->ge1 <= 0 && (int8_t)(b->go1 + b->ge1) <= 0 && (b->pw < 2 || (b->ge2 <= 0 && (int8_t)(b->go2 + b->ge2) <= 0))) ? 128 : 0;
	return (int64_t)bytes;
}

// ---- ingest / egress ---------------------------------------------------------------------------------------------
#include "io_host.cuh"

// ---- POA read-vs-graph sweep (bspoa.h:2515-2618) ---------------------------------------------------------------
#include "poa_host.cuh"
