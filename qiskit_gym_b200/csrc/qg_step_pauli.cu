// fused step kernel instantiations for one environment kind (see qg_step_kind.cuh)
#include "qg_step_kind.cuh"

namespace qg {
QG_INSTANTIATE_KIND(QG_ENV_PAULI_NETWORK)
}
