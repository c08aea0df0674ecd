// lintrans.h -- host-side views used by lintrans.cu (mirrors of lgpu_lintrans / lgpu_galois_keys of the public header).
#pragma once
#include <map>
#include <vector>
#include "composite.h"

namespace lgpu {

struct LinTransView {        // lintrans.LinearTransformation, circuits/common/lintrans/lintrans.go:150-160
    int level_q, level_p, log_slots, n1, n_diags;
    const int* diag_index;   // host: keys of Vec
    const u64* const* diag;  // host array of device pointers: (level_q+1) Q rows then (level_p+1) P rows, NTT + Montgomery
};
struct GaloisKeySet {        // the GaloisKeys map of rlwe.MemEvaluationKeySet, core/rlwe/evaluationkeyset.go
    int n = 0;
    const u64* gal_els = nullptr;
    std::vector<GadgetCt> keys;
    const GadgetCt* find(u64 galEl) const;   // nullptr + error "GaloisKey[..] is missing"
};

u64 galois_element(const Ctx* c, long long k);
void bsgs_index(const LinTransView& m, std::map<int, std::vector<int>>& index, std::vector<int>& rotN1, std::vector<int>& rotN2);
int automorphism_hoisted_lazy(const Ctx* c, int levelQ, CSpan ct0P, const u64* decomp, int decomp_levelQ, u64 galEl, const GadgetCt& gk,
                              const AccSpans& out, int batch, cudaStream_t st);
int lintrans_evaluate_many(const Ctx* c, int level_in, const u64* ct_in, const LinTransView* mats, int n_mats, const GaloisKeySet& gks,
                           u64* const* outs, int* out_levels, int batch, cudaStream_t st);

}  // namespace lgpu
