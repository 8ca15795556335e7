// l3ic bit-stream codec of the learned image codec on the device: replaces the host loop of compression/codec.py:87-265
// (scipy vq -> pyfse.compress per latent layer -> length table -> byte stream, and back) and pyfse's FSE_compress / FSE_decompress
// (pyfse/pyfse.pyx:24-72) for whole batches. Bit-exact with the reference library (tests/test_l3ic.py, oracle/_ref).
//
// Layout: a batch of n latents (n,h,w,c) float32 -> code-book indices, LAYER-major bytes (n,c,h*w) -> one coded layer per (image, layer)
// in a fixed-stride slot -> one byte stream per image in a fixed-stride slot:
//     [h w c : 3 x u8][len(coded lengths) : u16][coded lengths : FSE(u16[c]) or raw][layer 0][layer 1] ...
// with every layer FSE-coded, or 3 bytes (u16 count, u8 value) when it is one repeated symbol, or raw when FSE does not shrink it.
// Work mapping: entropy coding is serial inside a stream, so the parallelism is ACROSS streams — one warp per stream (thousands per
// launch), lanes cooperating on histogram / copies, lane 0 walking the state machine over shared-memory tables (csrc/fse_core.cuh).
#include "fse_core.cuh"
#include "ni_common.cuh"

namespace {

enum { kStatusOk = 0, kStatusFse = 1, kStatusShape = 2, kStatusTruncated = 3, kStatusSymbol = 4, kStatusLengths = 5, kStatusSingleByte = 6 };

constexpr int kStage = 4096;

__device__ __forceinline__ void flag(int* status, int img, int code) {
    if (status) atomicCAS(status + img, 0, code);
}

// values -> index of the nearest code-book entry (scipy.cluster.vq.vq: squared distance in double, first minimum wins); NHWC -> layer-major
__global__ void l3ic_quantise_kernel(const float* __restrict__ latent, const float* __restrict__ codebook, int n_codes, unsigned char* __restrict__ idx,
                                     long long total, int hw, int c) {
    extern __shared__ float cb[];
    for (int i = threadIdx.x; i < n_codes; i += blockDim.x) cb[i] = codebook[i];
    __syncthreads();
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const double v = (double)latent[e];
        int best = 0;
        double bd = (v - (double)cb[0]) * (v - (double)cb[0]);
        for (int k = 1; k < n_codes; ++k) {
            const double d = (v - (double)cb[k]) * (v - (double)cb[k]);
            if (d < bd) { bd = d; best = k; }
        }
        const long long pix = e / c;
        const int layer = (int)(e - pix * c);
        const long long img = pix / hw;
        const int p = (int)(pix - img * hw);
        idx[(img * c + layer) * hw + p] = (unsigned char)best;
    }
}

__global__ void l3ic_dequantise_kernel(const unsigned char* __restrict__ idx, int slot, const float* __restrict__ codebook, int n_codes,
                                       float* __restrict__ latent, long long total, int hw, int c, int* __restrict__ status) {
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const long long pix = e / c;
        const int layer = (int)(e - pix * c);
        const long long img = pix / hw;
        const int p = (int)(pix - img * hw);
        const int s = idx[(img * c + layer) * slot + p];
        if (s >= n_codes) { flag(status, (int)img, kStatusSymbol); latent[e] = 0.f; continue; }
        latent[e] = codebook[s];
    }
}

// One warp per stream. l3ic = 0: pyfse.compress semantics (dst_len = size | 0 not compressible | 1 repeated symbol | < 0 error).
// l3ic = 1: compression/codec.py:127-140 — the repeated-symbol and not-compressible cases fall back to the 3-byte run / the raw bytes.
__global__ void __launch_bounds__(32) fse_encode_kernel(const unsigned char* __restrict__ src, long long src_stride, const int* __restrict__ src_len,
                                                        int fixed_len, unsigned char* __restrict__ dst, long long dst_stride, int dst_cap,
                                                        int* __restrict__ dst_len, int l3ic, int streams_per_image, int* __restrict__ status) {
    __shared__ fse::EncScratch S;
    __shared__ int result;
    const int sid = blockIdx.x, lane = threadIdx.x;
    const unsigned char* in = src + (long long)sid * src_stride;
    unsigned char* out = dst + (long long)sid * dst_stride;
    const int n = src_len ? src_len[sid] : fixed_len;
    __shared__ unsigned char stage[kStage];                 // the serial walk of lane 0 reads its symbols from shared memory when they fit
    for (int s = lane; s < 256; s += 32) S.count[s] = 0;
    __syncwarp();
    const bool staged = n <= kStage;
    for (int i = lane; i < n; i += 32) {
        const unsigned char v = in[i];
        if (staged) stage[i] = v;
        atomicAdd(&S.count[v], 1u);
    }
    __syncwarp();
    if (staged) in = stage;
    if (lane == 0) result = n >= 0 ? fse::compress_counted(out, (uint32_t)dst_cap, in, (uint32_t)n, S) : fse::kErrSrcSizeWrong;
    __syncwarp();
    int r = result;
    if (l3ic) {
        const int img = sid / streams_per_image;
        if (r == 1) {                                   // np.uint16(len) + np.uint8(value)
            if (lane == 0) { out[0] = (unsigned char)(n & 0xFF); out[1] = (unsigned char)((n >> 8) & 0xFF); out[2] = in[0]; }
            r = 3;
        } else if (r == 0) {
            for (int i = lane; i < n && i < dst_cap; i += 32) out[i] = in[i];
            r = n;
        } else if (r < 0) {
            if (lane == 0) flag(status, img, kStatusFse);
        }
        if (r == 1 && lane == 0) flag(status, img, kStatusSingleByte);
    }
    if (lane == 0) dst_len[sid] = r;
}

// One CTA per image: length table, header, concatenation of the coded layers (compression/codec.py:160-186).
__global__ void __launch_bounds__(128) l3ic_pack_kernel(const unsigned char* __restrict__ layer_bytes, int layer_slot, const int* __restrict__ layer_len,
                                                        int h, int w, int c, unsigned char* __restrict__ streams, long long stream_stride,
                                                        int* __restrict__ stream_len, int* __restrict__ status) {
    __shared__ fse::EncScratch S;
    __shared__ unsigned char table[512], coded[512 + 16];
    __shared__ int offs[257];
    const int img = blockIdx.x, t = threadIdx.x;
    unsigned char* out = streams + (long long)img * stream_stride;
    const int* len = layer_len + (long long)img * c;
    if (t == 0) {
        for (int l = 0; l < c; ++l) { table[2 * l] = (unsigned char)(len[l] & 0xFF); table[2 * l + 1] = (unsigned char)((len[l] >> 8) & 0xFF); }
        fse::histogram(table, (uint32_t)(2 * c), S.count);
        int r = fse::compress_counted(coded, 512, table, (uint32_t)(2 * c), S);
        if (r == 1 || r < 0) { flag(status, img, kStatusLengths); r = 0; }      // pyfse raises (uncaught in the reference) for a constant table
        const unsigned char* lens = coded;
        if (r == 0) { lens = table; r = 2 * c; }
        out[0] = (unsigned char)h; out[1] = (unsigned char)w; out[2] = (unsigned char)c;
        out[3] = (unsigned char)(r & 0xFF); out[4] = (unsigned char)(r >> 8);
        for (int i = 0; i < r; ++i) out[5 + i] = lens[i];
        int o = 5 + r;
        for (int l = 0; l < c; ++l) { offs[l] = o; o += len[l] > 0 ? len[l] : 0; }
        offs[c] = o;
        stream_len[img] = o;
    }
    __syncthreads();
    for (int l = 0; l < c; ++l) {
        const unsigned char* in = layer_bytes + ((long long)img * c + l) * layer_slot;
        unsigned char* o = out + offs[l];
        const int n = offs[l + 1] - offs[l];
        for (int i = t; i < n; i += 128) o[i] = in[i];
    }
}

// One CTA per image, thread 0: header and length table of a stream (compression/codec.py:201-226) -> byte offset and size of every layer.
__global__ void __launch_bounds__(32) l3ic_parse_kernel(const unsigned char* __restrict__ streams, long long stream_stride, const int* __restrict__ stream_len,
                                                        int h, int w, int c, int* __restrict__ layer_off, int* __restrict__ layer_len,
                                                        int* __restrict__ status) {
    __shared__ fse::DecScratch S;
    __shared__ unsigned char lens[5120 + 16];
    if (threadIdx.x != 0) return;
    const int img = blockIdx.x;
    const unsigned char* in = streams + (long long)img * stream_stride;
    const int total = stream_len[img];
    int* off = layer_off + (long long)img * c;
    int* len = layer_len + (long long)img * c;
    for (int l = 0; l < c; ++l) { off[l] = 0; len[l] = -1; }
    if (total < 5) { flag(status, img, kStatusTruncated); return; }
    if (in[0] != (unsigned char)h || in[1] != (unsigned char)w || in[2] != (unsigned char)c) { flag(status, img, kStatusShape); return; }
    const int nl = in[3] | in[4] << 8;
    if (5 + nl > total) { flag(status, img, kStatusTruncated); return; }
    const unsigned char* tab = in + 5;
    if (nl > 2 * c) { flag(status, img, kStatusLengths); return; }      // a coded table is always shorter than the raw one (bounds `lens`)
    if (nl != 2 * c) {
        const int got = fse::decompress(lens, (uint32_t)(10 * nl), tab, (uint32_t)nl, S);      // pyfse.decompress default capacity: 10 x input
        if (got < 2 * c) { flag(status, img, kStatusLengths); return; }
        tab = lens;
    }
    int o = 5 + nl;
    for (int l = 0; l < c; ++l) {
        const int n = tab[2 * l] | tab[2 * l + 1] << 8;
        if (o + n > total) { flag(status, img, kStatusTruncated); return; }
        off[l] = o; len[l] = n; o += n;
    }
}

// One warp per stream. l3ic = 0: pyfse.decompress (dst_len = decoded size or < 0). l3ic = 1: compression/codec.py:243-255 — 3 bytes = run,
// h*w bytes = raw, else FSE; the layer must decode to exactly `expect` symbols.
__global__ void __launch_bounds__(32) fse_decode_kernel(const unsigned char* __restrict__ src, long long src_stride, const int* __restrict__ src_off,
                                                        const int* __restrict__ src_len, unsigned char* __restrict__ dst, long long dst_stride,
                                                        int dst_cap, int* __restrict__ dst_len, int l3ic, int expect, int streams_per_image,
                                                        int* __restrict__ status) {
    __shared__ fse::DecScratch S;
    __shared__ int result;
    const int sid = blockIdx.x, lane = threadIdx.x;
    const int img = l3ic ? sid / streams_per_image : sid;
    const unsigned char* in = src + (long long)(l3ic ? img : sid) * src_stride + (src_off ? src_off[sid] : 0);
    unsigned char* out = dst + (long long)sid * dst_stride;
    const int n = src_len[sid];
    int r;
    if (l3ic && n < 0) {
        r = -1;                                         // the parser already flagged this image
    } else if (l3ic && n == 3) {
        const int run = in[0] | in[1] << 8;
        for (int i = lane; i < run && i < dst_cap; i += 32) out[i] = in[2];
        r = run;
    } else if (l3ic && n == expect) {
        for (int i = lane; i < n; i += 32) out[i] = in[i];
        r = n;
    } else {
        __shared__ unsigned char stage[kStage];
        if (n > 0 && n <= kStage) {
            for (int i = lane; i < n; i += 32) stage[i] = in[i];
            in = stage;
        }
        __syncwarp();
        if (lane == 0) result = fse::decompress(out, (uint32_t)dst_cap, in, (uint32_t)(n > 0 ? n : 0), S);
        __syncwarp();
        r = result;
    }
    if (l3ic && r != expect && lane == 0) flag(status, img, r < 0 ? kStatusFse : kStatusSymbol);
    if (lane == 0 && dst_len) dst_len[sid] = r;
}

inline int grid_for_elems(long long total) {
    long long g = (total + 255) / 256;
    const long long cap = (long long)ni_num_sms() * 16;
    return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace

extern "C" int ni_fse_compress_batch(const unsigned char* src, long long src_stride, const int* src_len, unsigned char* dst, long long dst_stride,
                                     int* dst_len, int n, cudaStream_t st) {
    NI_REQUIRE(src && src_len && dst && dst_len && n > 0 && src_stride > 0 && dst_stride > 0, "ni_fse_compress_batch: invalid arguments");
    fse_encode_kernel<<<n, 32, 0, st>>>(src, src_stride, src_len, 0, dst, dst_stride, (int)(dst_stride > 0x7fffffff ? 0x7fffffff : dst_stride), dst_len, 0,
                                        1, nullptr);
    NI_LAUNCH_CHECK();
    NI_COUNT_LAUNCH(1);
    return NI_OK;
}

extern "C" int ni_fse_decompress_batch(const unsigned char* src, long long src_stride, const int* src_len, unsigned char* dst, long long dst_stride,
                                       int dst_cap, int* dst_len, int n, cudaStream_t st) {
    NI_REQUIRE(src && src_len && dst && dst_len && n > 0 && src_stride > 0 && dst_cap > 0 && dst_stride >= dst_cap,
               "ni_fse_decompress_batch: invalid arguments");
    fse_decode_kernel<<<n, 32, 0, st>>>(src, src_stride, nullptr, src_len, dst, dst_stride, dst_cap, dst_len, 0, 0, 1, nullptr);
    NI_LAUNCH_CHECK();
    NI_COUNT_LAUNCH(1);
    return NI_OK;
}

extern "C" int ni_l3ic_encode(const float* latent, int n, int h, int w, int c, const float* codebook, int n_codes, unsigned char* indices,
                              unsigned char* layer_bytes, int layer_slot, int* layer_len, unsigned char* streams, long long stream_stride,
                              int* stream_len, int* status, cudaStream_t st) {
    NI_REQUIRE(latent && codebook && indices && layer_bytes && layer_len && streams && stream_len && status, "ni_l3ic_encode: null pointer");
    NI_REQUIRE(n > 0 && h > 0 && w > 0 && c > 0 && h <= 255 && w <= 255 && c <= 255, "ni_l3ic_encode: latent shape must fit three bytes (got %d x %d x %d)", h, w, c);
    NI_REQUIRE(n_codes >= 1 && n_codes <= 256, "ni_l3ic_encode: code-books with more than 256 centers are not supported");
    const int hw = h * w;
    NI_REQUIRE(hw >= 2 && hw <= 65535, "ni_l3ic_encode: h * w must lie in [2, 65535]");
    NI_REQUIRE(layer_slot >= hw && stream_stride >= 5 + 2LL * c + (long long)c * hw, "ni_l3ic_encode: output slots too small");
    const long long total = (long long)n * hw * c;
    NI_CUDA(cudaMemsetAsync(status, 0, sizeof(int) * n, st));
    l3ic_quantise_kernel<<<grid_for_elems(total), 256, sizeof(float) * n_codes, st>>>(latent, codebook, n_codes, indices, total, hw, c);
    NI_LAUNCH_CHECK();
    fse_encode_kernel<<<n * c, 32, 0, st>>>(indices, hw, nullptr, hw, layer_bytes, layer_slot, layer_slot, layer_len, 1, c, status);
    NI_LAUNCH_CHECK();
    l3ic_pack_kernel<<<n, 128, 0, st>>>(layer_bytes, layer_slot, layer_len, h, w, c, streams, stream_stride, stream_len, status);
    NI_LAUNCH_CHECK();
    NI_COUNT_LAUNCH(3);
    return NI_OK;
}

extern "C" int ni_l3ic_decode(const unsigned char* streams, long long stream_stride, const int* stream_len, int n, int h, int w, int c,
                              const float* codebook, int n_codes, int* layer_off, int* layer_len, unsigned char* indices, int index_slot,
                              float* latent, int* status, cudaStream_t st) {
    NI_REQUIRE(streams && stream_len && codebook && layer_off && layer_len && indices && latent && status, "ni_l3ic_decode: null pointer");
    NI_REQUIRE(n > 0 && h > 0 && w > 0 && c > 0 && h <= 255 && w <= 255 && c <= 255, "ni_l3ic_decode: latent shape must fit three bytes (got %d x %d x %d)", h, w, c);
    NI_REQUIRE(n_codes >= 1 && n_codes <= 256, "ni_l3ic_decode: code-books with more than 256 centers are not supported");
    const int hw = h * w;
    NI_REQUIRE(index_slot >= hw + 4, "ni_l3ic_decode: index_slot must be at least h * w + 4");
    const long long total = (long long)n * hw * c;
    NI_CUDA(cudaMemsetAsync(status, 0, sizeof(int) * n, st));
    l3ic_parse_kernel<<<n, 32, 0, st>>>(streams, stream_stride, stream_len, h, w, c, layer_off, layer_len, status);
    NI_LAUNCH_CHECK();
    fse_decode_kernel<<<n * c, 32, 0, st>>>(streams, stream_stride, layer_off, layer_len, indices, index_slot, index_slot, nullptr, 1, hw, c, status);
    NI_LAUNCH_CHECK();
    l3ic_dequantise_kernel<<<grid_for_elems(total), 256, 0, st>>>(indices, index_slot, codebook, n_codes, latent, total, hw, c, status);
    NI_LAUNCH_CHECK();
    NI_COUNT_LAUNCH(3);
    return NI_OK;
}
