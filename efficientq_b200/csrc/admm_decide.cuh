// Best-iterate bookkeeping of one scored ADMM iterate (reference EfficientQConv.py:118-122, :139-142),
// called by ONE thread: by admm_decide_kernel after a conv-scored iterate, and by the tail of
// quadform_delta_kernel, which produces the squared error itself.
#pragma once
#include "common.cuh"

namespace effq {

__device__ __forceinline__ void admm_decide_dev(effq_admm_state* st, double total_sse, double numel, float* history) {
  const float loss = (float)(total_sse / numel);             // F.mse_loss(...).item()
  const int it = st->iter;
  const int keep = (it == 0 || loss < st->best_loss) ? 1 : 0;   // strict <, iterate 0 always seeds (:139)
  if (keep) { st->best_loss = loss; st->best_iter = it; st->best_conv_scale = st->conv_scale; }
  st->last_loss = loss;
  st->sse = total_sse;
  if (history) history[it] = loss;
  st->iter = it + 1;
  st->take_ = keep;
}

}  // namespace effq
