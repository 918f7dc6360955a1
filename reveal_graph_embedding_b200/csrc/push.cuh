// push.cuh -- parameters shared by the two walk schedules of the push engine:
//   push.cu           exact FIFO replay of the reference's queue discipline (default)
//   push_frontier.cu  synchronous frontier rounds on fixed-point state (opt-in, tolerance parity)
#pragma once

#include "common.cuh"

namespace arcte {

struct PushParams {
    int64_t n;
    const NodeInfo *info;      // {d_in, row begin, row length} per node
    const int32_t *indices;
    const double *w;
    const double2 *wd;         // {w, d_in of the target} per stored entry (FIFO schedule)
    // work list
    const int32_t *work_seed;  // [n_work_total] seed node per position
    const double *work_eps;    // [n_work_total]
    const int32_t *work_ids;   // positions to run (retry pass) or nullptr = 0..n_work-1
    int64_t n_work;
    int retry_pass;
    // slots
    double2 *sr;
    int32_t *touched;
    int32_t *queue;
    int64_t queue_cap;  // power of two
    int64_t n_slots;
    // outputs
    int32_t *seg_count;
    int64_t *seg_offset;
    int32_t *members;
    int64_t member_cap;
    int32_t *retry_list;
    unsigned long long *counters;
    // rule constants, computed on the host exactly as Python evaluates them
    double rho;            // rho
    double one_minus_rho;  // (1-rho)
    double lazy_b;         // (1-rho)*(1-lazy)
    double lazy_c;         // (1-rho)*lazy
    int debug_keep;        // operator seam: stop after the walk, leave s/r in slot 0
    // frontier schedule only (push_frontier.cu)
    int32_t *frontier;     // [n_slots][2][n] current / next frontier
    double *fval;          // [n_slots][n] residual mass taken from each frontier node this round
    double scale;          // fixed-point unit: 2^F
    double inv_scale;      // 2^-F
    int64_t work_lo;       // first work-list position of this launch (work_ids == nullptr)
    int cursor;            // PushCounter index of the work cursor this launch pulls from
};

// Implemented in push_frontier.cu.
int frontier_plan_slots(arcte_cuda_ctx *c, int64_t n_work, int64_t *n_slots);
int frontier_ensure_slots(arcte_cuda_ctx *c, int64_t n_slots);
int frontier_launch(arcte_cuda_ctx *c, PushParams P, int64_t n_work, bool retry_pass);
double frontier_scale(double rho);

}  // namespace arcte
