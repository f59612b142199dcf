// sweeper_factory_b200.cpp -- repo-local replacement of the ONE reference translation unit
// that has to change for a new sweeper type to be selectable from the XML input:
// src/sweepers/transport_sweeper_factory.cpp:28-70 (a hard-coded string switch). Same entry
// point and behaviour (mocc::TransportSweeperFactory, declared in
// src/sweepers/transport_sweeper_factory.hpp), plus
//   <sweeper type="moc_cuda" ...>       the B200 MoC sweeper (CudaMoCSweeper)
//   <sweeper type="2d3d_cuda" ...>      2D3D with the B200 MoC sweeper on every plane
// All reference types ("moc", "sn", "2d3d", "moc_2d3d") still build the reference classes.
#include "sweepers/transport_sweeper_factory.hpp"

#include <functional>
#include <map>
#include <string>

#include "core/mesh.hpp"
#include "sweepers/cmdo/plane_sweeper_2d3d.hpp"
#include "sweepers/moc/moc_sweeper.hpp"
#include "sweepers/sn_sweeper_factory.hpp"
#include "util/error.hpp"
#include "util/files.hpp"

#include "cuda_moc_sweeper.hpp"

namespace mocc_b200 {
// plane_sweeper_2d3d_cuda.cpp: the reference's PlaneSweeper_2D3D compiled around the CUDA MoC sweeper
mocc::UP_Sweeper_t make_plane_sweeper_2d3d_cuda(const pugi::xml_node &input, const mocc::CoreMesh &mesh);
}

namespace mocc {
namespace {
using Maker = std::function<UP_Sweeper_t(const pugi::xml_node &, const CoreMesh &)>;

const std::map<std::string, std::pair<const char *, Maker>> &registry()
{
    static const std::map<std::string, std::pair<const char *, Maker>> reg = {
        {"moc",
         {"Using an MoC sweeper",
          [](const pugi::xml_node &n, const CoreMesh &m) { return UP_Sweeper_t(new moc::MoCSweeper(n, m)); }}},
        {"moc_cuda",
         {"Using the B200 (CUDA) MoC sweeper",
          [](const pugi::xml_node &n, const CoreMesh &m) {
              return UP_Sweeper_t(new mocc_b200::CudaMoCSweeper(n, m));
          }}},
        {"sn",
         {"Using an Sn sweeper",
          [](const pugi::xml_node &n, const CoreMesh &m) { return UP_Sweeper_t(SnSweeperFactory(n, m)); }}},
        {"2d3d",
         {"Using a 2D3D sweeper",
          [](const pugi::xml_node &n, const CoreMesh &m) {
              return UP_Sweeper_t(new cmdo::PlaneSweeper_2D3D(n, m));
          }}},
        {"2d3d_cuda",
         {"Using a 2D3D sweeper with the B200 (CUDA) MoC sweeper",
          [](const pugi::xml_node &n, const CoreMesh &m) {
              return mocc_b200::make_plane_sweeper_2d3d_cuda(n, m);
          }}},
        {"moc_2d3d_cuda",
         {"Using a standalone 2D3D B200 (CUDA) MoC sweeper",
          [](const pugi::xml_node &n, const CoreMesh &m) {
              auto *swp = new mocc_b200::CudaMoCSweeper2D3D(n, m);
              swp->set_self_coupling();
              return UP_Sweeper_t(swp);
          }}},
        {"moc_2d3d",
         {"Using a standalone 2D3D MoC sweeper",
          [](const pugi::xml_node &n, const CoreMesh &m) {
              // only useful for one-way coupling, as in the reference
              auto *swp = new cmdo::MoCSweeper_2D3D(n, m);
              swp->set_self_coupling();
              return UP_Sweeper_t(swp);
          }}},
    };
    return reg;
}
}

UP_Sweeper_t TransportSweeperFactory(const pugi::xml_node &input, const CoreMesh &mesh)
{
    LogFile << "Generating transport sweeper..." << std::endl;
    const pugi::xml_node node = input.child("sweeper");
    const std::string type    = node.attribute("type").value();
    const auto it             = registry().find(type);
    if (it == registry().end())
        throw EXCEPT("Failed to detect a valid sweeper type.");
    LogScreen << it->second.first << std::endl;
    UP_Sweeper_t sweeper = it->second.second(node, mesh);
    LogFile << "Done generating transport sweeper." << std::endl;
    return sweeper;
}
}
