// mocc_flatten -- XML input -> .mocflat: the flattened ray-tracing data the C ABI consumes
// (flatten.hpp) plus, with --xs, the per-FSR macroscopic cross sections of every group.
// This is the host-side setup step of the plugin run stand-alone, so that non-C++ hosts
// (mocc_b200/capi.py, bench.py) can drive the same C ABI on the same problem.
//
//   mocc_flatten <in.xml> <out.mocflat> [--xs] [--set path/to/node@attr=value]...
#include <cstdio>
#include <iostream>
#include <string>
#include <vector>

#include "pugixml.hpp"

#include "core/core_mesh.hpp"
#include "sweepers/moc/moc_sweeper.hpp"
#include "util/error.hpp"
#include "util/files.hpp"
#include "util/timers.hpp"

#include "arrayfile.hpp"
#include "flatten.hpp"
#include "xml_amend.hpp"

using namespace mocc;

namespace {
// Read access to the ray data and cross-section mesh a MoCSweeper builds
class Probe : public moc::MoCSweeper {
public:
    Probe(const pugi::xml_node &input, const CoreMesh &mesh) : moc::MoCSweeper(input, mesh)
    {
    }
    mocc_b200::FlatProblem flat() const
    {
        std::vector<double> vol(vol_.begin(), vol_.end());
        return mocc_b200::flatten(mesh_, rays_, macroplane_unique_ids_, first_reg_macroplane_, vol.data(),
                                  (int)n_reg_, (int)n_group_);
    }
    void put_xs(mocc_b200::ArrayFile &af) const
    {
        const int ng = (int)n_group_, nr = (int)n_reg_;
        std::vector<double> tr((size_t)ng * nr), self(tr.size()), nf(tr.size()), ch(tr.size());
        std::vector<double> scat((size_t)ng * ng * nr, 0.0); // [to][from][reg]
        for (const auto &xsr : *xs_mesh_) {
            for (int g = 0; g < ng; g++) {
                const ScatteringRow &row = xsr.xsmacsc().to(g);
                for (const int r : xsr.reg()) {
                    const size_t o = (size_t)g * nr + r;
                    tr[o] = xsr.xsmactr(g), nf[o] = xsr.xsmacnf(g), ch[o] = xsr.xsmacch(g);
                    self[o] = row[g];
                    for (int gf = row.min_g; gf <= row.max_g; gf++)
                        scat[((size_t)g * ng + gf) * nr + r] = row[gf];
                }
            }
        }
        af.put("xs_tr", tr.data(), {(uint64_t)ng, (uint64_t)nr});
        af.put("xs_self", self.data(), {(uint64_t)ng, (uint64_t)nr});
        af.put("xs_nf", nf.data(), {(uint64_t)ng, (uint64_t)nr});
        af.put("xs_ch", ch.data(), {(uint64_t)ng, (uint64_t)nr});
        af.put("xs_scat", scat.data(), {(uint64_t)ng, (uint64_t)ng, (uint64_t)nr});
        af.put_scalar<int32_t>("n_inner", (int32_t)n_inner_);
        af.put_scalar<int32_t>("gs_boundary", gauss_seidel_boundary_ ? 1 : 0);
    }
};
}

int main(int argc, char **argv)
{
    try {
        if (argc < 3) {
            std::cerr << "usage: mocc_flatten <in.xml> <out.mocflat> [--xs] [--set path@attr=value]...\n";
            return 2;
        }
        bool xs = false;
        std::vector<std::string> sets;
        for (int i = 3; i < argc; i++) {
            const std::string a = argv[i];
            if (a == "--xs")
                xs = true;
            else if (a == "--set" && i + 1 < argc)
                sets.push_back(argv[++i]);
            else
                throw std::runtime_error("unknown argument: " + a);
        }
        pugi::xml_document doc;
        const auto res = doc.load_file(argv[1]);
        if (!res)
            throw std::runtime_error(std::string("cannot parse ") + argv[1] + ": " + res.description());
        for (const auto &s : sets)
            mocc_b200::amend_xml(doc, s);
        StartLogFile("mocc_flatten");
        RootTimer.tic();
        CoreMesh mesh(doc);
        Probe sw(doc.child("solver").child("sweeper"), mesh);
        mocc_b200::FlatProblem fp = sw.flat();
        mocc_b200::ArrayFile af   = fp.to_arrayfile();
        if (xs)
            sw.put_xs(af);
        af.save(argv[2]);
        std::printf("mocc_flatten: %lld segments (%lld resident), %lld rays, %d FSRs, %d groups, %d planes\n",
                    (long long)fp.n_seg_reference, (long long)fp.seg_len.size(), (long long)fp.n_ray_reference,
                    fp.n_reg, fp.n_group, fp.n_plane);
        return 0;
    } catch (const mocc::Exception &e) {
        std::cerr << "mocc_flatten: " << e.what() << std::endl;
        return 1;
    } catch (const std::exception &e) {
        std::cerr << "mocc_flatten: " << e.what() << std::endl;
        return 1;
    }
}
