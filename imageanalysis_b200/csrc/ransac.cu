// ransac.cu — placeholder until the batched 5-point kernel lands (see ransac.h).
#include "ransac.h"

namespace iam {

int ransac_pairs(int, const float*, const float*, const int32_t*, int, const double*, double, double, int, uint32_t,
                 uint8_t*, double*, int32_t*, cudaStream_t, std::string* err) {
  if (err) *err = "RANSAC kernels not built in this revision";
  return -5;  // IAM_E_UNSUPPORTED
}

}  // namespace iam
