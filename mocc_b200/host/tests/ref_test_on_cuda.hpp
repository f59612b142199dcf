// ref_test_on_cuda.hpp -- forced in front of the reference's OWN unit tests of the MoC sweeper
// (src/sweepers/moc/tests/test_MoC_IHM.cpp, test_MoCSweeper.cpp; compiled where they lie, unmodified) so that
// every `MoCSweeper` they name is the CUDA sweeper of this repository ("implication (1)").
//   test_MoC_IHM     infinite homogeneous medium, 7 groups, 800 inners per group: flux within 0.5 % of the
//                    analytic spectrum (test_MoC_IHM.cpp:136-147) -- through CudaMoCSweeper::sweep
//   test_MoCSweeper  pin-flux get / set / get round trip (test_MoCSweeper.cpp:58-94) on the subclass
#pragma once
#include "sweepers/moc/moc_sweeper.hpp" // the real class first: the tests' own include of it becomes a no-op

#include "cuda_moc_sweeper.hpp"

namespace mocc {
namespace moc {
using MoCSweeperOnB200 = mocc_b200::CudaMoCSweeper;
}
}
#define MoCSweeper MoCSweeperOnB200
