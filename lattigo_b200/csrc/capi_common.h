// capi_common.h -- shared by the extern "C" translation units (capi.cu, capi2.cu, ringmisc.cu): the opaque handle,
// argument-check macros and the row-map helpers.
#pragma once
#include <cstdint>
#include "../../include/lattigo_b200.h"
#include "engine.h"
#include <nvtx3/nvToolsExt.h>

namespace lgpu {
struct NvtxRange {
    explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
    NvtxRange(const NvtxRange&) = delete;
    NvtxRange& operator=(const NvtxRange&) = delete;
};
}  // namespace lgpu

struct lgpu_ctx {
    lgpu::Ctx c;
};

#define REQUIRE(cond, msg)                 \
    do {                                   \
        if (!(cond)) {                     \
            lgpu::set_error(msg);          \
            return -1;                     \
        }                                  \
    } while (0)
// every device entry point opens an NVTX range named after itself (visible in nsys / ncu --nvtx; a no-op without a tool attached), so that a
// trace of the Go application shows which ring / evaluator method each group of kernels belongs to
#define REQUIRE_DEVICE(ctx)                                                                                                      \
    REQUIRE((ctx) && (ctx)->c.device >= 0, "this context was created host-only (device < 0): no device execution");             \
    lgpu::NvtxRange lgpu_nvtx_range_(__func__)
// the key-switch family moves 128 bits per access: polynomial blocks must be 16-byte aligned with even strides
#define AL(p) ((reinterpret_cast<uintptr_t>(p) & 15u) == 0)
#define REQUIRE_ALIGNED(cond) REQUIRE(cond, "polynomial buffers and evaluation keys must be 16-byte aligned with even strides (words)")

namespace lgpu {
// rows 0..level of ring Q or P -> global limb indices (capi.cu)
int make_rowmap(const Ctx& c, int ring, int level, RowMap& rm);
int make_rowmap_single(const Ctx& c, int ring, int limb, RowMap& rm);
}  // namespace lgpu
