// wire.cu -- the reference's binary wire format parsed straight into the device layout (evaluation keys
// and polynomials reach HBM without an intermediate Go / host object). Everything is little-endian uint64 words
// (utils/buffer/writer.go:311-330):
//   ring.Poly            = structs.Matrix[uint64]: rows, then per row: len, len words      ring/poly.go:132-179, utils/structs/matrix.go:80-140,
//                                                                                          utils/structs/vector.go:82-143
//   ringqp.Poly          = Q poly, P poly (0 rows when there is no P)                      ring/ringqp/poly.go:105-170
//   rlwe.VectorQP        = len, then len ringqp.Poly                                       core/rlwe/keys.go:153-189
//   rlwe.GadgetCiphertext= BaseTwoDecomposition, Matrix[VectorQP]: digits, per digit: pw2 count, per entry a VectorQP (len 2)
//                                                                                          core/rlwe/gadgetciphertext.go:101-167
//   rlwe.EvaluationKey / RelinearizationKey = the GadgetCiphertext; rlwe.GaloisKey = GaloisElement, NthRoot, EvaluationKey
//                                                                                          core/rlwe/keys.go:628-700
// Device layout produced: lgpu_gadget_ct.data = [digit][pw2][component][Q rows | P rows][N] (include/lattigo_b200.h).
#include <cstring>
#include <string>
#include <vector>
#include "capi_common.h"

using namespace lgpu;

namespace {

struct Reader {
    const unsigned char* p; size_t n, off;
    bool u64at(uint64_t& v) {
        if (off + 8 > n) return false;
        uint64_t x = 0;
        for (int i = 7; i >= 0; i--) x = (x << 8) | p[off + i];     // little endian, whatever the host is
        v = x; off += 8;
        return true;
    }
};

struct RowRef { size_t off; };   // byte offset of a row's first word

// ring.Poly: records where each row's words start; every row must hold exactly N words
static int parse_poly(Reader& r, size_t N, std::vector<RowRef>& rows, int& nrows) {
    uint64_t nr;
    if (!r.u64at(nr)) { set_error("wire: truncated ring.Poly header"); return -1; }
    if (nr > (uint64_t)kMaxRows) { set_error("wire: ring.Poly with too many rows"); return -1; }
    nrows = (int)nr;
    for (int i = 0; i < nrows; i++) {
        uint64_t len;
        if (!r.u64at(len)) { set_error("wire: truncated ring.Poly row header"); return -1; }
        if (len != N) { set_error("wire: ring.Poly row of " + std::to_string(len) + " coefficients, ring degree is " + std::to_string(N)); return -1; }
        if (r.off + len * 8 > r.n) { set_error("wire: truncated ring.Poly row"); return -1; }
        rows.push_back(RowRef{r.off});
        r.off += len * 8;
    }
    return 0;
}

static bool host_little_endian() { const uint16_t x = 1; return *reinterpret_cast<const unsigned char*>(&x) == 1; }

}  // namespace

extern "C" {

// ring.Poly.ReadFrom / UnmarshalBinary into `rows_cap` device rows of N words; *rows_out = rows found, *consumed = bytes read
int lgpu_poly_load(lgpu_ctx* ctx, const void* bytes, size_t nbytes, uint64_t* dst, int rows_cap, int* rows_out, size_t* consumed, void* stream) {
    REQUIRE_DEVICE(ctx);
    REQUIRE(bytes && dst, "null argument");
    REQUIRE(host_little_endian(), "wire loaders need a little-endian host");
    Reader r{(const unsigned char*)bytes, nbytes, 0};
    std::vector<RowRef> rows;
    int n = 0;
    if (parse_poly(r, ctx->c.N, rows, n)) return -1;
    REQUIRE(n <= rows_cap, "wire: destination polynomial has fewer rows than the encoded one");
    const size_t N = ctx->c.N;
    for (int i = 0; i < n; i++)
        LGPU_CUDA_OK(cudaMemcpyAsync(dst + (size_t)i * N, r.p + rows[i].off, N * 8, cudaMemcpyHostToDevice, (cudaStream_t)stream));
    if (rows_out) *rows_out = n;
    if (consumed) *consumed = r.off;
    return 0;
}

// ring.Poly.WriteTo / MarshalBinary of `rows` device rows; bytes needed = 8 + rows * (8 + 8 N). Synchronises the stream.
int lgpu_poly_store(lgpu_ctx* ctx, const uint64_t* src, int rows, void* bytes, size_t cap, size_t* written, void* stream) {
    REQUIRE_DEVICE(ctx);
    REQUIRE(src && bytes, "null argument");
    REQUIRE(host_little_endian(), "wire loaders need a little-endian host");
    const size_t N = ctx->c.N, need = 8 + (size_t)rows * (8 + 8 * N);
    REQUIRE(rows >= 0 && cap >= need, "wire: output buffer too small");
    unsigned char* o = (unsigned char*)bytes;
    uint64_t w = (uint64_t)rows;
    memcpy(o, &w, 8); o += 8;
    for (int i = 0; i < rows; i++) {
        w = N; memcpy(o, &w, 8); o += 8;
        LGPU_CUDA_OK(cudaMemcpyAsync(o, src + (size_t)i * N, N * 8, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
        o += N * 8;
    }
    LGPU_CUDA_OK(cudaStreamSynchronize((cudaStream_t)stream));
    if (written) *written = need;
    return 0;
}

// rlwe.GadgetCiphertext.ReadFrom. Pass dst == NULL to size the key: fills *info (levels, digits, pw2 sizes, device_bytes,
// consumed) without touching the device; then call again with a device block of info->device_bytes bytes (16-byte aligned).
int lgpu_gadget_ct_load(lgpu_ctx* ctx, const void* bytes, size_t nbytes, uint64_t* dst, size_t dst_bytes, lgpu_evk_info* info, void* stream) {
    REQUIRE(ctx && bytes && info, "null argument");
    REQUIRE(host_little_endian(), "wire loaders need a little-endian host");
    const size_t N = ctx->c.N;
    Reader r{(const unsigned char*)bytes, nbytes, 0};
    uint64_t pw2, nd;
    REQUIRE(r.u64at(pw2) && r.u64at(nd), "wire: truncated GadgetCiphertext header");
    REQUIRE(nd >= 1 && nd <= LGPU_MAX_DIGITS, "wire: GadgetCiphertext digit count out of range");
    struct Entry { std::vector<RowRef> q[2], p[2]; };
    std::vector<std::vector<Entry>> val(nd);
    int lq = -2, lp = -2, maxpw2 = 0;
    for (uint64_t i = 0; i < nd; i++) {
        uint64_t nj;
        REQUIRE(r.u64at(nj), "wire: truncated GadgetCiphertext digit header");
        REQUIRE(nj >= 1 && nj <= 64, "wire: BaseTwoDecompositionVectorSize out of range");
        val[i].resize(nj);
        info->pw2_sizes[i] = (int)nj;
        if ((int)nj > maxpw2) maxpw2 = (int)nj;
        for (uint64_t j = 0; j < nj; j++) {
            uint64_t deg;
            REQUIRE(r.u64at(deg), "wire: truncated VectorQP header");
            REQUIRE(deg == 2, "wire: GadgetCiphertext entries must be degree-1 (two ringqp.Poly)");
            for (int k = 0; k < 2; k++) {
                int nq = 0, np = 0;
                if (parse_poly(r, N, val[i][j].q[k], nq)) return -1;
                if (parse_poly(r, N, val[i][j].p[k], np)) return -1;
                if (lq == -2) { lq = nq - 1; lp = np - 1; }
                REQUIRE(nq - 1 == lq && np - 1 == lp, "wire: GadgetCiphertext entries have different levels");
            }
        }
    }
    REQUIRE(lq >= 0 && lq < ctx->c.nQ && lp >= -1 && lp < ctx->c.nP, "wire: GadgetCiphertext levels exceed this context's moduli chains");
    const size_t rows = (size_t)(lq + 1) + (size_t)(lp + 1);
    info->level_q = lq; info->level_p = lp; info->base_two_decomposition = (int)pw2; info->n_digits = (int)nd; info->n_pw2_max = maxpw2;
    info->device_bytes = (size_t)nd * maxpw2 * 2 * rows * N * 8;
    info->consumed = r.off;
    if (!dst) return 0;
    REQUIRE_DEVICE(ctx);
    REQUIRE_ALIGNED(AL(dst));
    REQUIRE(dst_bytes >= info->device_bytes, "wire: device block smaller than lgpu_evk_info.device_bytes");
    cudaStream_t st = (cudaStream_t)stream;
    for (uint64_t i = 0; i < nd; i++)
        for (size_t j = 0; j < val[i].size(); j++)
            for (int k = 0; k < 2; k++) {
                uint64_t* base = dst + ((((size_t)i * maxpw2 + j) * 2 + k) * rows) * N;
                for (int x = 0; x <= lq; x++)
                    LGPU_CUDA_OK(cudaMemcpyAsync(base + (size_t)x * N, r.p + val[i][j].q[k][x].off, N * 8, cudaMemcpyHostToDevice, st));
                for (int x = 0; x <= lp; x++)
                    LGPU_CUDA_OK(cudaMemcpyAsync(base + (size_t)(lq + 1 + x) * N, r.p + val[i][j].p[k][x].off, N * 8, cudaMemcpyHostToDevice, st));
            }
    return 0;
}

// rlwe.GaloisKey.ReadFrom (core/rlwe/keys.go:670-700): GaloisElement, NthRoot, then the EvaluationKey (= GadgetCiphertext)
int lgpu_galois_key_load(lgpu_ctx* ctx, const void* bytes, size_t nbytes, uint64_t* gal_el, uint64_t* nth_root, uint64_t* dst, size_t dst_bytes,
                         lgpu_evk_info* info, void* stream) {
    REQUIRE(ctx && bytes && info, "null argument");
    Reader r{(const unsigned char*)bytes, nbytes, 0};
    uint64_t g, nr;
    REQUIRE(r.u64at(g) && r.u64at(nr), "wire: truncated GaloisKey header");
    REQUIRE(nr == ctx->c.nthroot, "wire: GaloisKey.NthRoot does not match this ring");
    if (gal_el) *gal_el = g;
    if (nth_root) *nth_root = nr;
    const int rc = lgpu_gadget_ct_load(ctx, (const unsigned char*)bytes + 16, nbytes - 16, dst, dst_bytes, info, stream);
    if (!rc) info->consumed += 16;
    return rc;
}

}  // extern "C"
