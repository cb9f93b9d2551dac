// qg_policy_host.hpp — the policy handle behind qg_policy_create (shared by qg_policy.cu and the fused-search launch in qg_engine.cu).
#pragma once
#include <vector>

#include "qg_policy_kernels.cuh"

struct qg_policy {
    int device = 0;
    qg::PolicyDev d{};
    std::vector<float*> bufs;
    size_t smem = 0;
    long long* acc0 = nullptr;       // first-layer accumulators of the one-launch search, [CTAs][8][width[0]] (allocated on first use)
    size_t acc0_ctas = 0;
};
