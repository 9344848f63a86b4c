// placeholder: specialised kernel lands in the next commit
#include "pt_kernels.cuh"
namespace pt {
bool window45_supported(const WinArgs &) { return false; }
cudaError_t launch_window45(const WinArgs &, int, int, cudaStream_t) { return cudaErrorNotSupported; }
const char *window45_name() { return "dog_window45_argmax"; }
}
