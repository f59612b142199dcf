// plane_sweeper_2d3d_cuda.cpp -- the reference's 2D3D plane sweeper with the B200 MoC sweeper
// on every plane, WITHOUT copying or editing it.
//
// cmdo::PlaneSweeper_2D3D holds its MoC sweeper by value (`MoCSweeper_2D3D moc_sweeper_`,
// src/sweepers/cmdo/plane_sweeper_2d3d.hpp:246), so the type cannot be swapped at run time. This
// translation unit compiles the UNMODIFIED reference sources of that class a second time, from
// where they lie, under a type substitution done by the preprocessor:
//     PlaneSweeper_2D3D  ->  PlaneSweeper_2D3D_Cuda
//     MoCSweeper_2D3D    ->  mocc_b200::CudaMoCSweeper2D3D   (same interface, CUDA sweep)
// Everything else of the 2D3D method (Sn CDD sweep, transverse leakage, projection, residual
// bookkeeping) is the reference's code, unchanged. The factory reaches the new class through
// mocc_b200::make_plane_sweeper_2d3d_cuda() so that no other translation unit sees the macros.

// 1. the real CPU class first, under its own name (its header is `#pragma once`: never re-read)
#include "sweepers/cmdo/moc_sweeper_2d3d.hpp"
// 2. the CUDA class with the same interface
#include "cuda_moc_sweeper.hpp"
namespace mocc {
namespace cmdo {
using CudaMoCSweeper2D3D_t = mocc_b200::CudaMoCSweeper2D3D;
}
}
// 3. the reference's plane sweeper, header and implementation, with the two names substituted
#define PlaneSweeper_2D3D PlaneSweeper_2D3D_Cuda
#define MoCSweeper_2D3D CudaMoCSweeper2D3D_t
#include "sweepers/cmdo/plane_sweeper_2d3d.cpp"
#undef PlaneSweeper_2D3D
#undef MoCSweeper_2D3D

namespace mocc_b200 {
mocc::UP_Sweeper_t make_plane_sweeper_2d3d_cuda(const pugi::xml_node &input, const mocc::CoreMesh &mesh)
{
    return mocc::UP_Sweeper_t(new mocc::cmdo::PlaneSweeper_2D3D_Cuda(input, mesh));
}
}
