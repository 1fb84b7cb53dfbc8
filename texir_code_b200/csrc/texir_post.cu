// texir_post.cu -- IrT post-processing on the device: nearest-texel padding of the empty texels of
// an irradiance texture (tools/padding_texture.py:49-80) and the Radiance RGBE pixel codec behind the
// reference's cv2.imwrite / cv2.imread of .hdr files (trainer/generate_ir_texture.py:80-82,
// models/mat_nvdiffrast.py:73,89).  The run-length file format around the codec is host C++ at the end.
//
// Padding is an exact Euclidean feature transform in two passes over a [H,W] int32 scratch:
//   k_pad_columns_a/b  one thread per (column, row segment), coalesced across a warp: nearest non-empty row above / below
//   k_pad_rows      one CTA per row, the row's sites in shared memory: lower envelope (one thread; W steps of
//                   integer work, ~0.1 ms for 4096), then every thread binary-searches its texels and copies the
//                   source texel.  Two launches, a few ms for 4096^2; it runs once per generated texture.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "../../include/texir_b200.h"
#include "texir_host.h"
#include "texir_post.h"

using namespace texir;
using texir_host::fail;
using texir_host::finish_launch;

#define TXP_CUDA(call)                                                                  \
    do {                                                                                \
        cudaError_t e_ = (call);                                                        \
        if (e_ != cudaSuccess) return fail(std::string(#call) + ": " + cudaGetErrorString(e_), 2); \
    } while (0)

// Column pass, split into row segments so that W x S threads work instead of W (blockIdx.y = segment):
//   k_pad_columns_a  within the segment, downwards: nearest non-empty row at or above each texel (-1: none in the
//                    segment so far); per (segment, column) the first and the last non-empty row of the segment
//   k_pad_columns_b  carries those across segments and sweeps upwards: nearest of "above" and "below", ties -> above
// Rows are taken eight at a time so that eight rows' loads are in flight per thread.
__global__ void __launch_bounds__(32) k_pad_columns_a(const float* __restrict__ img, int H, int W, int C, int rows_per_seg,
                                                      int32_t* __restrict__ near_row, int32_t* __restrict__ seg_first,
                                                      int32_t* __restrict__ seg_last) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= W) return;
    const int r0 = blockIdx.y * rows_per_seg, r1 = min(H, r0 + rows_per_seg);
    int last = -1, first = -1;
    for (int i0 = r0; i0 < r1; i0 += 8) {
        bool empty[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int i = i0 + k < r1 ? i0 + k : r1 - 1;
            const float* px = img + ((int64_t)i * W + j) * C;
            float s = __ldg(px);
            for (int c = 1; c < C; ++c) s = __fadd_rn(s, __ldg(px + c));  // pad_texel_empty
            empty[k] = s == 0.0f;
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            if (i0 + k >= r1) break;
            if (!empty[k]) {
                last = i0 + k;
                if (first < 0) first = last;
            }
            near_row[(int64_t)(i0 + k) * W + j] = last;
        }
    }
    seg_first[(int64_t)blockIdx.y * W + j] = first;
    seg_last[(int64_t)blockIdx.y * W + j] = last;
}

__global__ void __launch_bounds__(32) k_pad_columns_b(int H, int W, int rows_per_seg, int segs,
                                                      const int32_t* __restrict__ seg_first,
                                                      const int32_t* __restrict__ seg_last,
                                                      int32_t* __restrict__ near_row) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= W) return;
    const int seg = blockIdx.y;
    const int r0 = seg * rows_per_seg, r1 = min(H, r0 + rows_per_seg);
    int carry_up = -1, next = -1;
    for (int s = seg - 1; s >= 0 && carry_up < 0; --s) carry_up = seg_last[(int64_t)s * W + j];
    for (int s = seg + 1; s < segs && next < 0; ++s) next = seg_first[(int64_t)s * W + j];
    for (int i0 = r1 - 1; i0 >= r0; i0 -= 8) {
        int up[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) up[k] = near_row[(int64_t)(i0 - k >= r0 ? i0 - k : r0) * W + j];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int i = i0 - k;
            if (i < r0) break;
            const int u = up[k] >= 0 ? up[k] : carry_up;
            if (u == i) next = i;
            near_row[(int64_t)i * W + j] = pad_pick_row(i, u, next);
        }
    }
}

// one warp-sized CTA per row: the row's sites as 16-bit values in shared memory (4 bytes per column, so 14 rows are
// resident per SM at W = 4096), lane 0 builds the envelope, all lanes query it
__global__ void __launch_bounds__(64) k_pad_rows(const float* __restrict__ img, int H, int W, int C,
                                                 const int32_t* __restrict__ near_row, float* __restrict__ out,
                                                 int32_t* __restrict__ index) {
    extern __shared__ unsigned short sm_pad[];
    unsigned short* nr = sm_pad;      // nearest non-empty row of every column, for this row
    unsigned short* g = sm_pad + W;   // envelope: kept columns
    __shared__ int s_maxl;
    const int i = blockIdx.x;
    for (int x = threadIdx.x; x < W; x += blockDim.x) {
        const int r = near_row[(int64_t)i * W + x];
        nr[x] = r < 0 ? kPadNone : (unsigned short)r;
    }
    __syncthreads();
    if (threadIdx.x == 0) s_maxl = pad_build_envelope(nr, W, i, g);
    __syncthreads();
    const int maxl = s_maxl;
    for (int x = threadIdx.x; x < W; x += blockDim.x) {
        int sr = i, sc = x;        // nearest non-empty texel
        int rr = i, rc = x;        // texel the value is read from
        if ((int)nr[x] != i && maxl >= 0) {
            sc = pad_query_envelope(nr, g, maxl, i, x);
            sr = nr[sc];
            rr = pad_source_index(sr, H);
            rc = pad_source_index(sc, W);
        }
        const int64_t o = ((int64_t)i * W + x) * C;
        for (int c = 0; c < C; ++c)
            out[o + c] = (rr >= 0 && rc >= 0) ? __ldg(img + ((int64_t)rr * W + rc) * C + c) : 0.0f;
        if (index) {
            index[2 * ((int64_t)i * W + x) + 0] = sr;
            index[2 * ((int64_t)i * W + x) + 1] = sc;
        }
    }
}

// pixel codec: img [n,3] float32 in RGB (bgr = 0) or BGR (bgr = 1, cv2's order) <-> rgbe [n,4] bytes (file order)
__global__ void __launch_bounds__(256) k_rgbe_encode(const float* __restrict__ img, int64_t n, int bgr,
                                                     uchar4* __restrict__ rgbe) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float a = img[3 * i + 0], b = img[3 * i + 1], c = img[3 * i + 2];
    unsigned char q[4];
    rgbe_from_float(bgr ? c : a, b, bgr ? a : c, q);
    rgbe[i] = make_uchar4(q[0], q[1], q[2], q[3]);
}
__global__ void __launch_bounds__(256) k_rgbe_decode(const uchar4* __restrict__ rgbe, int64_t n, int bgr,
                                                     float* __restrict__ img) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uchar4 v = rgbe[i];
    const unsigned char q[4] = {v.x, v.y, v.z, v.w};
    float r, g, b;
    rgbe_to_float(q, &r, &g, &b);
    img[3 * i + 0] = bgr ? b : r;
    img[3 * i + 1] = g;
    img[3 * i + 2] = bgr ? r : b;
}

// ------------------------------------------------------------------------------------------------
// Radiance .hdr container on the host (OpenCV HdrEncoder / HdrDecoder + rgbe.cpp): text header, then per
// scanline either flat RGBE quadruplets or the "new" run-length form (2, 2, width hi, width lo, then the four
// byte planes of the scanline, each run-length coded).
// ------------------------------------------------------------------------------------------------
static void rle_plane(const unsigned char* data, int numbytes, std::vector<unsigned char>& out) {
    const int kMinRun = 4;
    int cur = 0;
    while (cur < numbytes) {
        int beg_run = cur, run_count = 0, old_run_count = 0;
        while (run_count < kMinRun && beg_run < numbytes) {  // next run of at least 4, if any
            beg_run += run_count;
            old_run_count = run_count;
            run_count = 1;
            while (beg_run + run_count < numbytes && run_count < 127 && data[beg_run] == data[beg_run + run_count])
                ++run_count;
        }
        if (old_run_count > 1 && old_run_count == beg_run - cur) {  // the bytes before it are one short run
            out.push_back((unsigned char)(128 + old_run_count));
            out.push_back(data[cur]);
            cur = beg_run;
        }
        while (cur < beg_run) {  // literal bytes up to the run
            int n = beg_run - cur;
            if (n > 128) n = 128;
            out.push_back((unsigned char)n);
            out.insert(out.end(), data + cur, data + cur + n);
            cur += n;
        }
        if (run_count >= kMinRun) {
            out.push_back((unsigned char)(128 + run_count));
            out.push_back(data[beg_run]);
            cur += run_count;
        }
    }
}

static bool read_line(FILE* fp, std::string& line) {
    line.clear();
    int ch;
    while ((ch = fgetc(fp)) != EOF) {
        if (ch == '\n') return true;
        line.push_back((char)ch);
    }
    return !line.empty();
}

static int hdr_open(const char* path, FILE** fp_out, int* H, int* W) {
    FILE* fp = fopen(path, "rb");
    if (!fp) return fail(std::string("texir_hdr: cannot open ") + path);
    std::string line;
    bool format = false;
    while (read_line(fp, line)) {  // header lines up to the blank one
        if (line.empty()) break;
        if (line.rfind("FORMAT=32-bit_rle_rgbe", 0) == 0) format = true;
    }
    int h = 0, w = 0;
    if (!format || !read_line(fp, line) || sscanf(line.c_str(), "-Y %d +X %d", &h, &w) != 2 || h <= 0 || w <= 0) {
        fclose(fp);
        return fail(std::string("texir_hdr: not a '-Y h +X w' 32-bit_rle_rgbe file: ") + path);
    }
    *H = h;
    *W = w;
    *fp_out = fp;
    return 0;
}

extern "C" {

int texir_pad_texture(const float* img, int H, int W, int C, float* out, int32_t* index, int32_t* scratch,
                      void* stream) {
    if (H <= 0 || W <= 0 || C <= 0 || C > 4) return fail("texir_pad_texture: bad shape");
    if (!img || !out || !scratch) return fail("texir_pad_texture: NULL argument");
    if (img == out) return fail("texir_pad_texture: in-place padding is not supported");
    if (H >= 32768 || W >= 32768) return fail("texir_pad_texture: sides must be below 32768 texels");
    const size_t smem = 2 * sizeof(unsigned short) * (size_t)W;
    texir_host::ScopedDevice guard(texir_host::device_of(out));
    if (smem > 48 * 1024)
        TXP_CUDA(cudaFuncSetAttribute(k_pad_rows, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int segs = std::max(1, std::min(32, (H + 63) / 64)), rows_per_seg = (H + segs - 1) / segs;
    int32_t* seg_ends = nullptr;  // [2][segs][W]: first / last non-empty row of every (segment, column)
    TXP_CUDA(cudaMallocAsync((void**)&seg_ends, 2 * sizeof(int32_t) * (size_t)segs * W, (cudaStream_t)stream));
    const dim3 cgrid((W + 31) / 32, segs);
    k_pad_columns_a<<<cgrid, 32, 0, (cudaStream_t)stream>>>(img, H, W, C, rows_per_seg, scratch, seg_ends,
                                                            seg_ends + (size_t)segs * W);
    int rc = finish_launch("k_pad_columns_a");
    if (!rc) {
        k_pad_columns_b<<<cgrid, 32, 0, (cudaStream_t)stream>>>(H, W, rows_per_seg, segs, seg_ends,
                                                                seg_ends + (size_t)segs * W, scratch);
        rc = finish_launch("k_pad_columns_b");
    }
    cudaFreeAsync(seg_ends, (cudaStream_t)stream);
    if (rc) return rc;
    k_pad_rows<<<H, 64, smem, (cudaStream_t)stream>>>(img, H, W, C, scratch, out, index);
    return finish_launch("k_pad_rows");
}

int texir_rgbe_encode(const float* img, int64_t n, int bgr, uint8_t* rgbe, void* stream) {
    if (n < 0) return fail("texir_rgbe_encode: bad size");
    if (n == 0) return 0;
    if (!img || !rgbe) return fail("texir_rgbe_encode: NULL argument");
    texir_host::ScopedDevice guard(texir_host::device_of(rgbe));
    k_rgbe_encode<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(img, n, bgr, (uchar4*)rgbe);
    return finish_launch("k_rgbe_encode");
}

int texir_rgbe_decode(const uint8_t* rgbe, int64_t n, int bgr, float* img, void* stream) {
    if (n < 0) return fail("texir_rgbe_decode: bad size");
    if (n == 0) return 0;
    if (!img || !rgbe) return fail("texir_rgbe_decode: NULL argument");
    texir_host::ScopedDevice guard(texir_host::device_of(img));
    k_rgbe_decode<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>((const uchar4*)rgbe, n, bgr, img);
    return finish_launch("k_rgbe_decode");
}

int texir_hdr_write_host(const char* path, const uint8_t* rgbe_host, int H, int W) {
    if (!path || !rgbe_host || H <= 0 || W <= 0) return fail("texir_hdr_write_host: bad argument");
    std::vector<unsigned char> out;
    char head[128];
    const int hl = snprintf(head, sizeof(head), "#?RADIANCE\nFORMAT=32-bit_rle_rgbe\n\n-Y %d +X %d\n", H, W);
    out.insert(out.end(), head, head + hl);
    if (W < 8 || W > 0x7fff) {  // run-length coding is not allowed: flat quadruplets
        out.insert(out.end(), rgbe_host, rgbe_host + (size_t)H * W * 4);
    } else {
        // scanlines are coded independently: bands of rows go to host threads, the pieces are appended in order
        const int workers = (int)std::max(1u, std::min(std::min(std::thread::hardware_concurrency(), 16u), (unsigned)((H + 63) / 64)));
        std::vector<std::vector<unsigned char>> piece((size_t)workers);
        auto band = [&](int w) {
            std::vector<unsigned char> plane((size_t)W * 4);
            std::vector<unsigned char>& dst = piece[(size_t)w];
            const int y0 = (int)((int64_t)H * w / workers), y1 = (int)((int64_t)H * (w + 1) / workers);
            for (int y = y0; y < y1; ++y) {
                const uint8_t* row = rgbe_host + (size_t)y * W * 4;
                const unsigned char mark[4] = {2, 2, (unsigned char)(W >> 8), (unsigned char)(W & 0xFF)};
                dst.insert(dst.end(), mark, mark + 4);
                for (int x = 0; x < W; ++x)
                    for (int c = 0; c < 4; ++c) plane[(size_t)c * W + x] = row[4 * x + c];
                for (int c = 0; c < 4; ++c) rle_plane(plane.data() + (size_t)c * W, W, dst);
            }
        };
        std::vector<std::thread> pool;
        for (int w = 1; w < workers; ++w) pool.emplace_back(band, w);
        band(0);
        for (auto& t : pool) t.join();
        for (const auto& pc : piece) out.insert(out.end(), pc.begin(), pc.end());
    }
    FILE* fp = fopen(path, "wb");
    if (!fp) return fail(std::string("texir_hdr_write_host: cannot create ") + path);
    const bool ok = fwrite(out.data(), 1, out.size(), fp) == out.size();
    if (fclose(fp) != 0 || !ok) return fail(std::string("texir_hdr_write_host: short write to ") + path);
    return 0;
}

int texir_hdr_read_info(const char* path, int* H, int* W) {
    if (!path || !H || !W) return fail("texir_hdr_read_info: NULL argument");
    FILE* fp = nullptr;
    if (int rc = hdr_open(path, &fp, H, W)) return rc;
    fclose(fp);
    return 0;
}

int texir_hdr_read_host(const char* path, uint8_t* rgbe_host, int H, int W) {
    if (!path || !rgbe_host) return fail("texir_hdr_read_host: NULL argument");
    FILE* fp = nullptr;
    int h = 0, w = 0;
    if (int rc = hdr_open(path, &fp, &h, &w)) return rc;
    struct Closer {
        FILE* f;
        ~Closer() { fclose(f); }
    } closer{fp};
    if (h != H || w != W) return fail("texir_hdr_read_host: size differs from texir_hdr_read_info");
    const std::string bad = std::string("texir_hdr_read_host: truncated or malformed pixel data in ") + path;
    // the pixel data is read in one piece and parsed from memory (one fread instead of one per run)
    const long at = ftell(fp);
    if (at < 0 || fseek(fp, 0, SEEK_END) != 0) return fail(bad);
    const long end = ftell(fp);
    if (end < at || fseek(fp, at, SEEK_SET) != 0) return fail(bad);
    std::vector<unsigned char> data((size_t)(end - at));
    if (!data.empty() && fread(data.data(), 1, data.size(), fp) != data.size()) return fail(bad);
    const unsigned char* q = data.data();
    const unsigned char* const qe = q + data.size();
    if (W < 8 || W > 0x7fff) {
        if ((size_t)(qe - q) < (size_t)H * W * 4) return fail(bad);
        memcpy(rgbe_host, q, (size_t)H * W * 4);
        return 0;
    }
    std::vector<unsigned char> plane((size_t)W * 4);
    for (int y = 0; y < H; ++y) {
        uint8_t* row = rgbe_host + (size_t)y * W * 4;
        if (qe - q < 4) return fail(bad);
        if (q[0] != 2 || q[1] != 2 || (q[2] & 0x80)) {  // not run-length coded: the rest is flat
            const size_t rest = (size_t)(H - y) * W * 4;
            if ((size_t)(qe - q) < rest) return fail(bad);
            memcpy(row, q, rest);
            return 0;
        }
        if (((int)q[2] << 8 | q[3]) != W) return fail(bad);
        q += 4;
        for (int c = 0; c < 4; ++c) {
            unsigned char* p = plane.data() + (size_t)c * W;
            int x = 0;
            while (x < W) {
                if (qe - q < 2) return fail(bad);
                const unsigned char b0 = q[0], b1 = q[1];
                q += 2;
                if (b0 > 128) {  // run
                    const int n = b0 - 128;
                    if (n > W - x) return fail(bad);
                    memset(p + x, b1, n);
                    x += n;
                } else {  // literal bytes
                    const int n = b0;
                    if (n == 0 || n > W - x || qe - q < n - 1) return fail(bad);
                    p[x++] = b1;
                    memcpy(p + x, q, (size_t)(n - 1));
                    q += n - 1;
                    x += n - 1;
                }
            }
        }
        for (int x = 0; x < W; ++x)
            for (int c = 0; c < 4; ++c) row[4 * x + c] = plane[(size_t)c * W + x];
    }
    return 0;
}

// PNG scanline filters (PNG 1.2 section 6: None / Sub / Up / Average / Paeth), undone in place row by row.
// `filtered` holds H rows of 1 + row_bytes bytes (the inflated IDAT stream of a non-interlaced image), `out` H rows of
// row_bytes; bpp = bytes per complete pixel (6 for the 16-bit RGB index texture).  Host memory only.
int texir_png_unfilter_host(const uint8_t* filtered, int H, int row_bytes, int bpp, uint8_t* out) {
    if (!filtered || !out || H <= 0 || row_bytes <= 0 || bpp <= 0 || bpp > 8) return fail("texir_png_unfilter_host: bad argument");
    std::vector<uint8_t> zero((size_t)row_bytes, 0);
    for (int y = 0; y < H; ++y) {
        const uint8_t* src = filtered + (size_t)y * (row_bytes + 1);
        const int type = src[0];
        ++src;
        uint8_t* cur = out + (size_t)y * row_bytes;
        const uint8_t* up = y ? cur - row_bytes : zero.data();
        switch (type) {
        case 0: memcpy(cur, src, (size_t)row_bytes); break;
        case 1:
            for (int i = 0; i < row_bytes; ++i) cur[i] = (uint8_t)(src[i] + (i >= bpp ? cur[i - bpp] : 0));
            break;
        case 2:
            for (int i = 0; i < row_bytes; ++i) cur[i] = (uint8_t)(src[i] + up[i]);
            break;
        case 3:
            for (int i = 0; i < row_bytes; ++i) cur[i] = (uint8_t)(src[i] + (((i >= bpp ? cur[i - bpp] : 0) + up[i]) >> 1));
            break;
        case 4:
            for (int i = 0; i < row_bytes; ++i) {
                const int a = i >= bpp ? cur[i - bpp] : 0, b = up[i], c = i >= bpp ? up[i - bpp] : 0;
                const int pa = abs(b - c), pb = abs(a - c), pc = abs(a + b - 2 * c);
                cur[i] = (uint8_t)(src[i] + (pa <= pb && pa <= pc ? a : pb <= pc ? b : c));
            }
            break;
        default: return fail("texir_png_unfilter_host: unknown filter type in scanline");
        }
    }
    return 0;
}

}  // extern "C"
