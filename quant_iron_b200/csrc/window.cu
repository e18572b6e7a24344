// window.cu -- fused register-window executor (placeholder until the kernel lands).
#include "common.cuh"

namespace qi {

bool window_supported(const qi_state* s) { (void)s; return false; }

int run_circuit_windowed(qi_state* s, const std::vector<PhysGate>& gates) {
    for (const PhysGate& g : gates) QI_TRY(launch_simple_gate(s, g));
    return QI_OK;
}

}  // namespace qi
