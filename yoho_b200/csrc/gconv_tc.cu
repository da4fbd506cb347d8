// tcgen05 split-BF16 group convolution (placeholder until the tensor-core kernel lands in this file).
#include <vector>
#include "common.cuh"

int gconv_tc_pack(yoho_ctx*, GLayer&, const std::vector<float>&) { return YOHO_OK; }
bool gconv_tc_eligible(const GLayer&, const GConvArgs&) { return false; }
int gconv_tc_forward(yoho_ctx*, const GLayer&, const GConvArgs&, cudaStream_t) {
    yoho_set_error("tcgen05 group convolution not built");
    return YOHO_ERR_ARG;
}
