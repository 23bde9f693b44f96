// pcm_codec.cuh -- the HBM-bound kernels either side of the FIR:
//   pcm_decode_kernel : AudioSamples::readAll()            (reference ProcessFile.cp:40-41)
//   peak_abs_kernel   : VectorMath::max_mag() loop         (reference ProcessFile.cp:92-96)
//   pcm_encode_kernel : AudioSamples::normalize + writeAll (reference ProcessFile.cp:100,117)
//   synth_pcm_kernel  : counter-based synthetic PCM, mirrors
//                       oracle_synth_sample() operation for operation.
// Conventions (decision D2): x = int / 2^(bits-1);
// q = clamp(rint(y * scale * 2^(bits-1))), ties to even, no dither.
//
// Both codec kernels move the interleaved bytes with 128-bit coalesced global
// accesses through a shared-memory tile and touch the planar FP64 side with
// lane == consecutive frame, so every global transaction is a full sector.
#pragma once
#include "ptx_sm100.cuh"

namespace firgpu {

constexpr int CODEC_NT = 256;           // default threads per CTA (128 and 256 are supported)
constexpr int CODEC_TILE_BYTES = 0;     // interleaved bytes staged per tile; 0 = CODEC_TILE_SAMPLES samples
constexpr int CODEC_TILE_SAMPLES = 4096; // default tile: 32 KB on the planar FP64 side whatever the PCM format
constexpr int CODEC_SMEM_MAX = 72 * 1024;
constexpr int CODEC_CARVEOUT = 50;      // default shared-memory carveout of the codec kernels, percent
// Resident CTAs per SM the encoder's register budget is sized for.  The 24-bit instantiations fit 48
// registers without a spill: 5 CTAs instead of 4 hide more of the planar loads' latency (measured
// 0.86 -> 0.93 of HBM on 400 MB of stereo, 0.85 -> 0.92 on 8 channels).  The 16- and 32-bit ones
// need 64 -- forced down to 48 they spill 100 bytes per thread and lose a third of their speed.
template <int BITS>
struct EncodeBounds {
	static constexpr int MIN_CTAS = BITS == 24 ? 5 : 0; // 0 = unspecified: the compiler settles on 64 registers, no spill
};
constexpr int CODEC_UNROLL = 4;         // independent 128-bit global loads a thread keeps in flight

// Geometry of one launch.  A CTA is persistent: it walks tiles blockIdx.x, +gridDim.x, ...; the
// grid is the number of CTAs resident at once (asked of the runtime at launch).
// Frames per tile are a multiple of 2*nt when the frame is narrow enough, so that in the planar
// pass a "row" (nt threads x 2 consecutive frames) never straddles two channels and rows can be
// unrolled with their loads issued back to back; otherwise a multiple of 32 (generic loop).
struct CodecGeom {
	int frames;      // frames per tile
	int nt;          // threads per CTA
	unsigned smem;   // dynamic shared memory per CTA
};

__host__ __device__ inline uint32_t pad_byte(uint32_t b);

// The default tile holds CODEC_TILE_SAMPLES samples: the planar side (8 B a sample) is what the
// HBM time goes to, and a launch over a 100 MB file must still give every resident CTA a dozen
// tiles, or the CTAs that draw the last ones finish a tile-time after the others.
__host__ inline CodecGeom codec_geom(int fb, int channels, int tile_bytes = CODEC_TILE_BYTES, int nt = CODEC_NT)
{
	CodecGeom g;
	int f = tile_bytes > 0 ? tile_bytes / fb : CODEC_TILE_SAMPLES / channels;
	if (f < 2 * nt && nt > 128 && f >= 256) nt = 128; // wide frames: rows of 128 thread-pairs still fit
	g.nt = nt;
	if (f >= 2 * nt) f = f / (2 * nt) * (2 * nt);
	else {
		f &= ~31;
		if (f < 32) f = 32;
	}
	g.frames = f;
	g.smem = pad_byte((uint32_t) (f * fb) + 64u) + 16u;
	return g;
}

// The shared tile keeps the interleaved bytes in a SKEWED layout: one pad word after
// every 32 (word w lives at w + w/32).  Lanes of the planar pass are 2*frame_bytes
// apart; for wide frames (16 ch x 32 bit = 64 B) that is a multiple of 128 B and would
// put all 32 lanes on one bank -- with the skew every power-of-two stride is
// conflict-free, and so are the four word stores of a 16-byte chunk per lane.
__host__ __device__ inline uint32_t pad_word(uint32_t w) { return w + (w >> 5); }
__host__ __device__ inline uint32_t pad_byte(uint32_t b) { return b + ((b >> 7) << 2); }

// 4 bytes at an arbitrary byte offset of the skewed tile, packed little-endian
// (byte at `off` in bits 0..7).
__device__ __forceinline__ uint32_t lds_unaligned_u32(const uint32_t* tile, uint32_t off)
{
	const uint32_t w = off >> 2;
	const uint32_t lo = tile[pad_word(w)];
	const uint32_t hi = tile[pad_word(w + 1)];
	return __funnelshift_r(lo, hi, (off & 3u) << 3);
}

template <int BITS, bool BE>
__device__ __forceinline__ int32_t pcm_to_int(uint32_t u)
{
	if (BITS == 16) {
		const uint32_t v = BE ? __byte_perm(u, 0, 0x4401) : (u & 0xffffu);
		return (int32_t) (int16_t) v;
	} else if (BITS == 24) {
		const uint32_t v = BE ? __byte_perm(u, 0, 0x4012) : (u & 0xffffffu);
		return ((int32_t) (v << 8)) >> 8;
	} else {
		return (int32_t) (BE ? __byte_perm(u, 0, 0x0123) : u);
	}
}

template <int BITS, bool BE>
__device__ __forceinline__ uint32_t int_to_pcm(int32_t q)
{
	const uint32_t u = (uint32_t) q;
	if (BITS == 16) return BE ? __byte_perm(u, 0, 0x4401) : (u & 0xffffu);
	if (BITS == 24) return BE ? __byte_perm(u, 0, 0x4012) : (u & 0xffffffu);
	return BE ? __byte_perm(u, 0, 0x0123) : u;
}

// NB bytes of u (little-endian order, already endian-arranged) at logical byte offset
// b of the skewed tile, with the widest stores the alignment allows (an aligned
// 2- or 4-byte piece never straddles a pad).
template <int NB>
__device__ __forceinline__ void store_pcm_bytes(unsigned char* tile, uint32_t b, uint32_t u)
{
	if (NB == 2) {
		if (!(b & 1)) *reinterpret_cast<uint16_t*>(tile + pad_byte(b)) = (uint16_t) u;
		else {
			tile[pad_byte(b)] = (unsigned char) u;
			tile[pad_byte(b + 1)] = (unsigned char) (u >> 8);
		}
	} else if (NB == 4) {
		if (!(b & 3)) *reinterpret_cast<uint32_t*>(tile + pad_byte(b)) = u;
		else if (!(b & 1)) {
			*reinterpret_cast<uint16_t*>(tile + pad_byte(b)) = (uint16_t) u;
			*reinterpret_cast<uint16_t*>(tile + pad_byte(b + 2)) = (uint16_t) (u >> 16);
		} else {
#pragma unroll
			for (int k = 0; k < 4; ++k) tile[pad_byte(b + k)] = (unsigned char) (u >> (8 * k));
		}
	} else {
		if (!(b & 1)) {
			*reinterpret_cast<uint16_t*>(tile + pad_byte(b)) = (uint16_t) u;
			tile[pad_byte(b + 2)] = (unsigned char) (u >> 16);
		} else {
			tile[pad_byte(b)] = (unsigned char) u;
			*reinterpret_cast<uint16_t*>(tile + pad_byte(b + 1)) = (uint16_t) (u >> 8);
		}
	}
}

// The same when the caller knows (once per tile, uniformly for the CTA) that every sample offset
// is a multiple of the sample size: one store, no alignment tests.  Three-byte samples are never
// "aligned"; they keep the two-store form above.
template <int NB, bool ALIGNED>
__device__ __forceinline__ void store_pcm_at(unsigned char* tile, uint32_t b, uint32_t u)
{
	if (ALIGNED && NB == 2) *reinterpret_cast<uint16_t*>(tile + pad_byte(b)) = (uint16_t) u;
	else if (ALIGNED && NB == 4) *reinterpret_cast<uint32_t*>(tile + pad_byte(b)) = u;
	else store_pcm_bytes<NB>(tile, b, u);
}

// The sample at logical byte offset b of the skewed tile, in bits 0..8*NB-1 (upper bits unspecified).
template <int NB, bool ALIGNED>
__device__ __forceinline__ uint32_t load_pcm_at(const unsigned char* tile, uint32_t b)
{
	if (ALIGNED && NB == 2) return *reinterpret_cast<const uint16_t*>(tile + pad_byte(b));
	if (ALIGNED && NB == 4) return *reinterpret_cast<const uint32_t*>(tile + pad_byte(b));
	return lds_unaligned_u32(reinterpret_cast<const uint32_t*>(tile), b);
}

// 16 bytes that are read exactly once: bypass L1 allocation (the in-flight loads of a streaming
// kernel must not depend on how much of the SM's array is left to L1 once the tile is carved out).
__device__ __forceinline__ double2 ldg_stream_f64x2(const double* p)
{
	double2 v;
	asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
	return v;
}

__device__ __forceinline__ uint4 ldg_stream_u32x4(const void* p)
{
	uint4 v;
	asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
	             : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
	             : "l"(p));
	return v;
}

// Dynamic tile scheduler of the persistent codec CTAs: tiles are handed out by one global counter,
// so the CTAs finish together whatever n_tiles / gridDim is (a static stride leaves up to one tile
// per CTA of imbalance -- 10 % of a 100 us launch).  Every CTA fetches until it is told "no more":
// n_tiles + gridDim fetches in all, and whoever draws the last one puts the counter back to zero
// for the next launch on the stream.  Thread 0 fetches (into a shared slot) right after a tile's
// first barrier, so that the atomic's latency hides under the second half of the tile; everybody
// reads the slot after the tile's last barrier.
__device__ __forceinline__ long long fetch_tile(unsigned long long* counter, long long n_tiles)
{
	const unsigned long long t = atomicAdd(counter, 1ull);
	if (t == (unsigned long long) n_tiles + gridDim.x - 1) *counter = 0ull; // the last fetch of the launch
	return (long long) t;
}

template <bool B>
struct BoolTag {
	static constexpr bool value = B;
};

// Decode a window of the interleaved PCM into the zero-padded planar layout.
//   pcm            : first byte of logical frame avail_lo
//   [avail_lo, avail_hi) : logical frames present in pcm (real data)
//   g0             : logical frame that lands at x index 0  (= first output frame - H)
//   n_x            : x entries to write per channel (the whole pitch: zeros
//                    wherever the logical frame is not available)
//   F              : frames per tile (CodecGeom::frames)
// Persistent CTAs: tile = blockIdx.x, +gridDim.x, ...  Per tile: FILL (interleaved bytes ->
// skewed shared tile, CODEC_UNROLL independent LDG.128 per thread in flight), then the PLANAR
// pass (two consecutive frames of one channel per thread -> one STG.128).
template <int BITS, bool BE>
__global__ void __launch_bounds__(256)
pcm_decode_kernel(const unsigned char* __restrict__ pcm, long long avail_lo, long long avail_hi,
                  long long g0, long long n_x, int channels, double* __restrict__ x, long long x_pitch, int F,
                  unsigned long long* __restrict__ tile_counter)
{
	constexpr int NB = BITS / 8;
	constexpr int U = CODEC_UNROLL;
	extern __shared__ __align__(16) unsigned char tile[];
	__shared__ long long tile_slot;
	const int nt = blockDim.x, tid = threadIdx.x;
	const int fb = channels * NB;
	const long long n_tiles = (n_x + F - 1) / F;
	const double inv = 1.0 / (double) (1ll << (BITS - 1));
	const uint32_t* tw = reinterpret_cast<const uint32_t*>(tile);
	const bool rows_ok = (F % (2 * nt)) == 0;

	if (tid == 0) tile_slot = fetch_tile(tile_counter, n_tiles);
	__syncthreads();
	for (long long tl = tile_slot; tl < n_tiles; tl = tile_slot) {
		const long long i0 = tl * F;
		// logical frames of this tile that exist in pcm
		long long ga = g0 + i0, gb = ga + F;
		if (gb > g0 + n_x) gb = g0 + n_x;
		if (ga < avail_lo) ga = avail_lo;
		if (gb > avail_hi) gb = avail_hi;
		uint32_t mis = 0;
		if (gb > ga) {
			const unsigned char* p0 = pcm + (size_t) (ga - avail_lo) * fb;
			const uint32_t nbytes = (uint32_t) (gb - ga) * fb;
			mis = (uint32_t) (reinterpret_cast<uintptr_t>(p0) & 15u);
			const unsigned char* base = p0 - mis; // 16-byte aligned
			const uint32_t end = mis + nbytes;
			const uint32_t nvec = (end + 15u) >> 4;
			for (uint32_t v0 = tid; v0 < nvec; v0 += U * nt) {
				uint4 q[U];
				bool full[U];
#pragma unroll
				for (int k = 0; k < U; ++k) { // all the loads first: U x 16 B in flight per thread
					const uint32_t v = v0 + k * nt, lo = v << 4;
					full[k] = v < nvec && lo >= mis && lo + 16 <= end;
					if (full[k]) q[k] = ldg_stream_u32x4(base + lo); // 128-bit coalesced, read once
				}
#pragma unroll
				for (int k = 0; k < U; ++k) {
					const uint32_t v = v0 + k * nt, lo = v << 4, hi = lo + 16;
					if (full[k]) {
						uint32_t* d = reinterpret_cast<uint32_t*>(tile) + pad_word(v << 2);
						d[0] = q[k].x;
						d[1] = q[k].y;
						d[2] = q[k].z;
						d[3] = q[k].w;
					} else if (v < nvec) { // ragged head / tail: never touch bytes outside the payload
						const uint32_t a = lo > mis ? lo : mis, b = hi < end ? hi : end;
						for (uint32_t s = lo; s < hi; ++s) tile[pad_byte(s)] = (s >= a && s < b) ? base[s] : 0;
					}
				}
			}
		}
		__syncthreads();
		if (tid == 0) tile_slot = fetch_tile(tile_counter, n_tiles); // the next tile; read after the barrier below

		const int la = (int) (ga - g0 - i0), lb = (int) (gb - g0 - i0); // tile-local frames present in pcm
		int nloc = F;
		if (i0 + nloc > n_x) nloc = (int) (n_x - i0);
		if (rows_ok && nloc == F && la == 0 && lb == F) {
			// whole tile, all of it real data: rows of nt thread-pairs (one channel per row, so the
			// channel and the alignment class are uniform), no per-sample predicates, no division
			const int rows_per_ch = F / (2 * nt);
			auto planar = [&](auto tag) {
				constexpr bool AL = decltype(tag)::value;
				for (int c = 0; c < channels; ++c) {
					double* xc = x + (long long) c * x_pitch + i0;
					const uint32_t ob = mis + (uint32_t) c * NB;
#pragma unroll 2
					for (int j = 0; j < rows_per_ch; ++j) {
						const int fl = 2 * (j * nt + tid);
						const uint32_t o = ob + (uint32_t) fl * fb;
						const double v0 = (double) pcm_to_int<BITS, BE>(load_pcm_at<NB, AL>(tile, o)) * inv;
						const double v1 = (double) pcm_to_int<BITS, BE>(load_pcm_at<NB, AL>(tile, o + fb)) * inv;
						*reinterpret_cast<double2*>(xc + fl) = make_double2(v0, v1);
					}
				}
			};
			if (NB != 3 && (mis % NB) == 0) planar(BoolTag<true>{});
			else planar(BoolTag<false>{});
		} else {
			// edge tiles (zero padding, ragged end) and very wide frames: channel by channel, a thread
			// converts two consecutive frames and stores them with one 128-bit write (i0 and x_pitch
			// are even, so the pair is 16-byte aligned)
			for (int c = 0; c < channels; ++c) {
				double* xc = x + (long long) c * x_pitch + i0;
				const uint32_t cbase = mis + (uint32_t) c * NB - (uint32_t) la * fb;
				for (int fl = 2 * tid; fl < nloc; fl += 2 * nt) {
					double v0 = 0.0, v1 = 0.0;
					if (fl >= la && fl < lb)
						v0 = (double) pcm_to_int<BITS, BE>(lds_unaligned_u32(tw, cbase + (uint32_t) fl * fb)) * inv;
					if (fl + 1 >= la && fl + 1 < lb)
						v1 = (double) pcm_to_int<BITS, BE>(lds_unaligned_u32(tw, cbase + (uint32_t) (fl + 1) * fb)) * inv;
					if (fl + 1 < nloc) *reinterpret_cast<double2*>(xc + fl) = make_double2(v0, v1);
					else xc[fl] = v0;
				}
			}
		}
		__syncthreads(); // the tile is refilled by the next trip; tile_slot holds the next tile
	}
}

// Planar FP64 -> interleaved PCM.  gain = scale * 2^(bits-1) (one rounding, on
// the host); the product y*gain is rounded to binary64, then to the nearest
// integer with ties to even (cvt.rni), after clamping to the signed range.
// Persistent CTAs like the decoder.  Per tile: the PLANAR pass (CODEC_UNROLL independent
// LDG.128 of the parked signal in flight per thread, then quantise and drop the bytes into
// the skewed tile), then the DRAIN (128-bit coalesced stores of the interleaved bytes).
// q = clamp(rint(y * gain)): the product is rounded to binary64, converted with round-to-nearest-
// even and saturation to int32 (one F2I), then clamped as an integer.  rint() is monotone and the
// limits are integers, so clamping after rounding equals clamping before it -- the FP64
// min/max pair of the textbook form costs ~16 instructions per sample on this machine, and these
// kernels are issue-bound, not HBM-bound, until that is gone.
template <int BITS, bool BE>
__device__ __forceinline__ uint32_t quantise(double v, double gain)
{
	int q;
	const double p = __dmul_rn(v, gain);
	asm("cvt.rni.sat.s32.f64 %0, %1;" : "=r"(q) : "d"(p));
	if (BITS < 32) {
		constexpr int LIM = (int) (1ll << (BITS < 32 ? BITS - 1 : 1));
		q = max(min(q, LIM - 1), -LIM);
	}
	return int_to_pcm<BITS, BE>(q);
}

template <int BITS, bool BE>
__global__ void __launch_bounds__(256, EncodeBounds<BITS>::MIN_CTAS)
pcm_encode_kernel(const double* __restrict__ y, long long y_pitch, long long frames, int channels,
                  double gain, unsigned char* __restrict__ pcm, int F, unsigned long long* __restrict__ tile_counter)
{
	constexpr int NB = BITS / 8;
	constexpr int U = CODEC_UNROLL;
	extern __shared__ __align__(16) unsigned char tile[];
	__shared__ long long tile_slot;
	const int nt = blockDim.x, tid = threadIdx.x;
	const int fb = channels * NB;
	const long long n_tiles = (frames + F - 1) / F;
	const bool rows_ok = (F % (2 * nt)) == 0;

	if (tid == 0) tile_slot = fetch_tile(tile_counter, n_tiles);
	__syncthreads();
	for (long long tl = tile_slot; tl < n_tiles; tl = tile_slot) {
		const long long f0 = tl * F;
		long long f1 = f0 + F;
		if (f1 > frames) f1 = frames;
		const int nf = (int) (f1 - f0);
		unsigned char* p0 = pcm + (size_t) f0 * fb;
		const uint32_t mis = (uint32_t) (reinterpret_cast<uintptr_t>(p0) & 15u);

		if (rows_ok && nf == F) {
			// rows of nt thread-pairs, one channel per row: (c, j) advance as CTA-uniform counters,
			// U rows at a time with all their loads issued before the first is used
			const int rows_per_ch = F / (2 * nt), rows = channels * rows_per_ch;
			auto planar = [&](auto tag) {
				constexpr bool AL = decltype(tag)::value;
				int c = 0, j = 0;
				for (int r0 = 0; r0 < rows; r0 += U) {
					double2 v[U];
					uint32_t o[U];
#pragma unroll
					for (int k = 0; k < U; ++k) {
						if (r0 + k < rows) {
							const int fl = 2 * (j * nt + tid);
							o[k] = mis + (uint32_t) c * NB + (uint32_t) fl * fb;
							v[k] = ldg_stream_f64x2(y + (long long) c * y_pitch + f0 + fl);
							if (++j == rows_per_ch) {
								j = 0;
								++c;
							}
						}
					}
#pragma unroll
					for (int k = 0; k < U; ++k) {
						if (r0 + k < rows) {
							store_pcm_at<NB, AL>(tile, o[k], quantise<BITS, BE>(v[k].x, gain));
							store_pcm_at<NB, AL>(tile, o[k] + fb, quantise<BITS, BE>(v[k].y, gain));
						}
					}
				}
			};
			if (NB != 3 && (mis % NB) == 0) planar(BoolTag<true>{});
			else planar(BoolTag<false>{});
		} else {
			// ragged last tile and very wide frames: channel by channel, two consecutive frames per
			// thread (one 128-bit load when both exist)
			for (int c = 0; c < channels; ++c) {
				const double* yc = y + (long long) c * y_pitch + f0;
				const uint32_t dc = mis + (uint32_t) c * NB;
				for (int fl = 2 * tid; fl < nf; fl += 2 * nt) {
					double v0, v1 = 0.0;
					const bool two = fl + 1 < nf;
					if (two) {
						const double2 v = *reinterpret_cast<const double2*>(yc + fl);
						v0 = v.x;
						v1 = v.y;
					} else {
						v0 = yc[fl];
					}
					store_pcm_bytes<NB>(tile, dc + (uint32_t) fl * fb, quantise<BITS, BE>(v0, gain));
					if (two) store_pcm_bytes<NB>(tile, dc + (uint32_t) (fl + 1) * fb, quantise<BITS, BE>(v1, gain));
				}
			}
		}
		__syncthreads();
		if (tid == 0) tile_slot = fetch_tile(tile_counter, n_tiles); // the next tile; read after the barrier below

		unsigned char* base = p0 - mis;
		const uint32_t end = mis + (uint32_t) nf * fb;
		const uint32_t nvec = (end + 15u) >> 4;
		for (uint32_t v = tid; v < nvec; v += nt) {
			const uint32_t lo = v << 4, hi = lo + 16;
			if (lo >= mis && hi <= end) {
				const uint32_t* q = reinterpret_cast<const uint32_t*>(tile) + pad_word(v << 2);
				*reinterpret_cast<uint4*>(base + lo) = make_uint4(q[0], q[1], q[2], q[3]); // 128-bit coalesced
			} else {
				const uint32_t a = lo > mis ? lo : mis, b = hi < end ? hi : end;
				for (uint32_t s = a; s < b; ++s) base[s] = tile[pad_byte(s)];
			}
		}
		__syncthreads(); // the tile is rewritten by the next trip; tile_slot holds the next tile
	}
}

// Stand-alone peak: grid (blocks, channels); 128-bit loads, warp shuffle max,
// one atomicMax per warp.
__global__ void __launch_bounds__(256)
peak_abs_kernel(const double* __restrict__ y, long long y_pitch, long long frames,
                unsigned long long* __restrict__ peak)
{
	const double* yc = y + (long long) blockIdx.y * y_pitch;
	const long long pairs = frames >> 1;
	double m = 0.0;
	for (long long p = (long long) blockIdx.x * blockDim.x + threadIdx.x; p < pairs;
	     p += (long long) gridDim.x * blockDim.x) {
		const double2 v = __ldg(reinterpret_cast<const double2*>(yc) + p);
		m = fmax(m, fmax(fabs(v.x), fabs(v.y)));
	}
	if ((frames & 1) && blockIdx.x == 0 && threadIdx.x == 0) m = fmax(m, fabs(yc[frames - 1]));
	m = warp_max(m);
	if ((threadIdx.x & 31) == 0 && m > 0.0) atomicMax(peak, (unsigned long long) __double_as_longlong(m));
}

// ---- synthetic PCM ----------------------------------------------------------

__device__ __forceinline__ unsigned long long splitmix64(unsigned long long z)
{
	z += 0x9E3779B97F4A7C15ull;
	z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
	z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
	return z ^ (z >> 31);
}

__device__ __forceinline__ double synth_tri(long long n, long long period)
{
	const long long p = n % period, half = period / 2;
	const long long v = p < half ? p : period - p;
	return __ddiv_rn((double) (4 * v - 2 * half), (double) (2 * half));
}

// Explicit _rn intrinsics: no FMA contraction, so the integers equal the
// oracle's (compiled with -ffp-contract=off) bit for bit.
__device__ __forceinline__ long long synth_sample(unsigned long long seed, int channel, long long frame,
                                                  int bits, long long rate, double gain)
{
	const double fs = (double) (1ll << (bits - 1));
	const unsigned long long r =
		splitmix64(seed ^ splitmix64(((unsigned long long) channel << 48) ^ (unsigned long long) frame));
	const double noise =
		__dsub_rn(__dmul_rn(__dmul_rn((double) (r >> 11), 1.0 / 9007199254740992.0), 2.0), 1.0);
	long long p_rumble = rate / 5;
	if (p_rumble < 2) p_rumble = 2;
	long long p_tone = rate / 1000;
	if (p_tone < 2) p_tone = 2;
	double v = __dadd_rn(0.05, __dmul_rn(0.2, synth_tri(frame + 17 * channel, p_rumble)));
	v = __dadd_rn(v, __dmul_rn(0.3, synth_tri(frame + 5 * channel, p_tone)));
	v = __dadd_rn(v, __dmul_rn(0.1, noise));
	v = __dmul_rn(v, gain);
	double q = rint(__dmul_rn(v, fs));
	if (q < -fs) q = -fs;
	if (q > fs - 1.0) q = fs - 1.0;
	return (long long) q;
}

__global__ void __launch_bounds__(256)
synth_pcm_kernel(unsigned long long seed, long long first_frame, long long frames, int channels, int bits,
                 int big_endian, long long rate, double gain, unsigned char* __restrict__ pcm)
{
	const int nb = bits / 8;
	const long long total = frames * channels;
	for (long long s = (long long) blockIdx.x * blockDim.x + threadIdx.x; s < total;
	     s += (long long) gridDim.x * blockDim.x) {
		const long long f = s / channels;
		const int c = (int) (s - f * channels);
		uint32_t u = (uint32_t) synth_sample(seed, c, first_frame + f, bits, rate, gain);
		unsigned char* d = pcm + (size_t) s * nb;
		if (big_endian)
			for (int b = nb - 1; b >= 0; --b) { d[b] = (unsigned char) u; u >>= 8; }
		else
			for (int b = 0; b < nb; ++b) { d[b] = (unsigned char) u; u >>= 8; }
	}
}

} // namespace firgpu
