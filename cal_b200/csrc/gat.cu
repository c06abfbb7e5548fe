// gat.cu -- GATConv backbone layers of CausalGAT (model.py:340, 388-390).
#include "internal.cuh"

namespace cal {

int launch_gat_forward(const Ctx& c, int layer, cudaStream_t s) {
  (void)c; (void)layer; (void)s;
  return CAL_EUNSUPPORTED;
}
int launch_gat_backward(const Ctx& c, int layer, cudaStream_t s) {
  (void)c; (void)layer; (void)s;
  return CAL_EUNSUPPORTED;
}

}  // namespace cal
