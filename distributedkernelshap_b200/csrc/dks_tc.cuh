// tcgen05 / TMEM version of the fused coalition kernel (placeholder until the tensor-core path lands).
#pragma once

#include "dks_common.cuh"

namespace dks {

inline bool tc_supported(const dks_ctx*, const ExplainParams&) { return false; }
inline int tc_launch(dks_ctx*, const ExplainParams&) { return DKS_ERR_UNSUPPORTED; }
inline int tc_fit(dks_ctx*) { return DKS_OK; }
inline int tc_plan_changed(dks_ctx*, int) { return DKS_OK; }
inline void tc_release(dks_ctx*) {}

}  // namespace dks
