// aqe_plan.hpp -- sample plans: the position lists of CustomBPlusDB's samplers in
// closed form.  A plan is either a short list of affine segments (expanded on the device by
// plan_position() in aqe_kernels.cuh, so no index array ever exists in HBM) or an explicit index list.
// Pure host arithmetic; nothing here touches CUDA.
#pragma once

#include <cstdint>
#include <string>
#include <vector>

#include "aqe_b200.h"

struct aqe_plan {
    std::vector<aqe_segment> segs;
    std::vector<uint64_t> seg_start;  // exclusive prefix of segs[].count, segs.size()+1 entries
    std::vector<int64_t> idx;         // explicit positions when segs is empty
    uint64_t count = 0;
    bool by_amount_order = false;     // positions index the amount-sorted permutation (stratified)
    uint64_t n_rows = 0;              // rows of the table the plan was built for (positions are < n_rows); 0: a caller-supplied list
    // device mirror, owned by the engine (aqe_engine.cu)
    void* d_segs = nullptr;
    void* d_start = nullptr;
    void* d_idx = nullptr;
    int device = -1;
};

namespace aqe {

// |cached_records_| as the reference's samplers see it after a load of n rows (custom_bplus_db.cpp:188-191).
uint64_t cache_rows(uint64_t n);
// Bulk-load shape of the reference B+ tree (sorted inserts, MAX_KEYS = 255).
uint64_t leaf_count(uint64_t n);
uint64_t tree_height(uint64_t n);

// Data the data-dependent samplers need from the device side.
struct PlanData {
    const double* zone_var = nullptr;  // adaptive_block: variance of each of the 10 zones
    int64_t clt_kstop = -1;            // clt_validated: lock-step stop step (0 = ran to completion)
    int64_t clt_stopper = -1;          //               thread index that raised should_stop
};

// Builds the plan.  Returns an aqe_status; `err` receives a message on failure.
int plan_build(uint64_t n_rows, int method, const aqe_sample_params& p, const PlanData& data, aqe_plan& out,
               std::string& err);
void plan_finalize(aqe_plan& pl);  // computes seg_start / count
int64_t plan_position_host(const aqe_plan& pl, uint64_t k);

// clt_validated_dual_pointer_sample: the per-thread stride sequences (custom_bplus_db.cpp:918-987).
struct CltThread { uint64_t first, step, len; int fast; };
int clt_threads(uint64_t n_rows, const aqe_sample_params& p, std::vector<CltThread>& out, int64_t& T, std::string& err);

// seeded stand-in for std::random_device draws: Philox(key=seed, ctr=(draw, 0, "SEED", method))
uint64_t seeded_u64(uint64_t seed, uint32_t method, uint64_t draw);
uint64_t seeded_below(uint64_t seed, uint32_t method, uint64_t draw, uint64_t bound);

}  // namespace aqe
