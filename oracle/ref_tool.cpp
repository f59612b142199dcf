// ref_tool.cpp -- TEST INFRASTRUCTURE ONLY (oracle side).
//
// Links the UNMODIFIED reference (oracle/_ref/libmocc_ref.a) and
//   flat    <in.xml> <out.mocflat>              flatten the reference's ray data
//   golden  <in.xml> <out_prefix>               flat file + per-sweep1g records produced by the
//                                               reference kernel (moc_sweeper_kernel.inc.hpp:36-180)
//   solve   <in.xml> <out.golden>               run the reference solver unchanged; dump k, flux
//   time    <in.xml> [--sweeps N] [--warmup W]  time N passes (every group once, n_inner inners each) of the
//                                               reference CPU sweep after W untimed passes (JSON to stdout)
// Options (any command):
//   --set path/to/node@attr=value               amend the XML before it is parsed
//   --outers N  --records "o:g:i,o:g:i,..."     (golden) outer iterations; which sweeps to record
//   --cmfd                                      (golden) attach a CoarseData so that the last
//                                               inner runs moc::Current
// Nothing in the product path (mocc_b200/) calls into this file.
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <iostream>
#include <memory>
#include <set>
#include <sstream>
#include <string>
#include <tuple>
#include <vector>

#include <omp.h>

#include "pugixml.hpp"

#include "core/coarse_data.hpp"
#include "core/core_mesh.hpp"
#include "core/source.hpp"
#include "solvers/eigen_solver.hpp"
#include "solvers/solver_factory.hpp"
#include "sweepers/moc/moc_current_worker.hpp"
#include "sweepers/moc/moc_sweeper.hpp"
// the 2D3D MoC sweeper keeps its coupling objects private; the tool records them
#define private protected
#include "sweepers/cmdo/moc_sweeper_2d3d.hpp"
#undef private
#include "sweepers/cmdo/correction_worker.hpp"
#include "util/files.hpp"

#include "arrayfile.hpp"
#include "flatten.hpp"

using namespace mocc;
using mocc_b200::ArrayFile;

namespace {

struct RecordKey {
    int outer, group, inner;
    bool operator<(const RecordKey &o) const
    {
        return std::tie(outer, group, inner) < std::tie(o.outer, o.group, o.inner);
    }
};

// Gives the tool access to the sweeper's protected state and to the reference kernel
// sweep1g<>. sweep_recorded() follows MoCSweeper::sweep (moc_sweeper.cpp:189-225) step by
// step so that inputs and outputs of every sweep1g call can be captured;
// Exposed2D3D::sweep_recorded() does the same for MoCSweeper_2D3D::sweep
// (cmdo/moc_sweeper_2d3d.cpp:48-99) with the reference's CurrentCorrections worker.
template <class Base> class ExposedT : public Base {
public:
    ExposedT(const pugi::xml_node &input, const CoreMesh &mesh) : Base(input, mesh)
    {
    }

    mocc_b200::FlatProblem flat() const
    {
        std::vector<double> vol(this->vol_.begin(), this->vol_.end());
        return mocc_b200::flatten(this->mesh_, this->rays_, this->macroplane_unique_ids_,
                                  this->first_reg_macroplane_, vol.data(), (int)this->n_reg_, (int)this->n_group_);
    }
    Source *source()
    {
        return this->source_;
    }
    int n_inner() const
    {
        return this->n_inner_;
    }
    bool gs() const
    {
        return this->gauss_seidel_boundary_;
    }
    int bc_per_group() const
    {
        return this->boundary_[0].size() / this->n_group_;
    }
    std::vector<double> bc_in(int group) const
    {
        std::vector<double> out;
        int n = bc_per_group();
        for (const auto &b : this->boundary_) {
            const real_t *p = b.get_boundary(group, 0).second;
            out.insert(out.end(), p, p + n);
        }
        return out;
    }
    std::vector<double> xs_per_reg(int group, int which) const
    {
        std::vector<double> out(this->n_reg_);
        for (const auto &xsr : *this->xs_mesh_) {
            double v = 0.0;
            switch (which) {
            case 0:
                v = xsr.xsmactr(group);
                break;
            case 1:
                v = xsr.xsmacsc().to(group)[group];
                break;
            case 2:
                v = xsr.xsmacnf(group);
                break;
            case 3:
                v = xsr.xsmacch(group);
                break;
            }
            for (const int ireg : xsr.reg())
                out[ireg] = v;
        }
        return out;
    }
    // scattering matrix expanded to FSRs: out[gfrom][ireg] for scattering INTO `group`
    std::vector<double> xs_scat_to(int group) const
    {
        std::vector<double> out((size_t)this->n_group_ * this->n_reg_, 0.0);
        for (const auto &xsr : *this->xs_mesh_) {
            const ScatteringRow &row = xsr.xsmacsc().to(group);
            for (int gf = row.min_g; gf <= row.max_g; gf++)
                for (const int ireg : xsr.reg())
                    out[(size_t)gf * this->n_reg_ + ireg] = row[gf];
        }
        return out;
    }

    void record_inputs(ArrayFile &out, const std::string &p, int outer, int group, int inner)
    {
        out.put_scalar<int32_t>(p + "outer", outer);
        out.put_scalar<int32_t>(p + "group", group);
        out.put_scalar<int32_t>(p + "inner", inner);
        std::vector<double> fin(this->flux_1g_.begin(), this->flux_1g_.end());
        out.put(p + "flux_in", fin);
        const VectorX &s = this->source_->get();
        out.put(p + "src", s.data(), {(uint64_t)s.size()});
        // every group's flux as Source::in_scatter saw it when it built `src` (the swept group's own column is not
        // read there): [n_reg][n_group]
        std::vector<double> fall(this->flux_.begin(), this->flux_.end());
        out.put(p + "flux_all", fall.data(), {(uint64_t)this->flux_.extent(0), (uint64_t)this->flux_.extent(1)});
        out.put(p + "bc_in", bc_in(group));
    }
    void record_source(ArrayFile &out, const std::string &p)
    {
        const VectorX &q = this->source_->get_transport(0);
        out.put(p + "qbar", q.data(), {(uint64_t)q.size()});
        std::vector<double> x(this->xstr_.xs().begin(), this->xstr_.xs().end());
        out.put(p + "xstr", x);
    }
    void record_outputs(ArrayFile &out, const std::string &p, int group, int mode)
    {
        out.put_scalar<int32_t>(p + "mode", mode);
        std::vector<double> fout(this->flux_1g_.begin(), this->flux_1g_.end());
        out.put(p + "flux_out", fout);
        out.put(p + "bc_out", bc_in(group));
        if (mode != 0) {
            auto all = blitz::Range::all();
            auto c   = this->coarse_data_->current(all, group);
            auto sf  = this->coarse_data_->surface_flux(all, group);
            std::vector<double> cv(c.begin(), c.end()), sv(sf.begin(), sf.end());
            out.put(p + "current", cv);
            out.put(p + "surface_flux", sv);
        }
    }

    virtual void sweep_recorded(int group, int outer, const std::set<RecordKey> &want, ArrayFile &out, int &n_rec)
    {
        this->xstr_.expand(group, this->split_);
        this->flux_1g_.reference(this->flux_(blitz::Range::all(), group));
        for (unsigned int inner = 0; inner < this->n_inner_; inner++) {
            bool rec = want.count(RecordKey{outer, group, (int)inner}) != 0;
            std::string p = "rec" + std::to_string(n_rec) + "_";
            if (rec)
                record_inputs(out, p, outer, group, (int)inner);
            this->source_->self_scatter(group, this->xstr_.xs());
            if (rec)
                record_source(out, p);
            int mode = 0;
            if (inner == this->n_inner_ - 1 && this->coarse_data_) {
                this->coarse_data_->zero_data_radial(group);
                moc::Current cw(this->coarse_data_, &this->mesh_);
                this->sweep1g(group, cw);
                this->coarse_data_->set_has_radial_data(true);
                mode = 1;
            } else {
                moc::NoCurrent cw(this->coarse_data_, &this->mesh_);
                this->sweep1g(group, cw);
            }
            if (rec) {
                record_outputs(out, p, group, mode);
                n_rec++;
            }
        }
    }
};
using ExposedMoC = ExposedT<moc::MoCSweeper>;

class Exposed2D3D : public ExposedT<cmdo::MoCSweeper_2D3D> {
public:
    Exposed2D3D(const pugi::xml_node &input, const CoreMesh &mesh) : ExposedT<cmdo::MoCSweeper_2D3D>(input, mesh)
    {
        set_self_coupling();
    }
    void sweep_recorded(int group, int outer, const std::set<RecordKey> &want, ArrayFile &out, int &n_rec) override
    {
        xstr_.expand(group, split_);
        if (allow_splitting_)
            xstr_true_.expand(group);
        cmdo::CurrentCorrections ccw(coarse_data_, &mesh_, corrections_.get(), source_->get_transport(0), xstr_true_,
                                     xstr_, xstr_sn_, ang_quad_, rays_);
        moc::NoCurrent ncw(coarse_data_, &mesh_);
        flux_1g_.reference(flux_(blitz::Range::all(), group));
        for (unsigned int inner = 0; inner < n_inner_; inner++) {
            bool rec = want.count(RecordKey{outer, group, (int)inner}) != 0;
            std::string p = "rec" + std::to_string(n_rec) + "_";
            if (rec)
                record_inputs(out, p, outer, group, (int)inner);
            source_->self_scatter(group, xstr_.xs());
            if (rec)
                record_source(out, p);
            int mode = 0;
            if (inner == n_inner_ - 1 && coarse_data_) {
                coarse_data_->zero_data_radial(group);
                sn_xs_mesh_->update();
                this->sweep1g(group, ccw);
                coarse_data_->set_has_radial_data(true);
                mode = 2;
            } else {
                this->sweep1g(group, ncw);
            }
            if (rec) {
                record_outputs(out, p, group, mode);
                if (mode == 2) {
                    std::vector<double> xt(xstr_true_.xs().begin(), xstr_true_.xs().end());
                    out.put(p + "xstr_true", xt);
                    // homogenised XS the beta factor divides by (correction_worker.cpp:84-93)
                    auto fp = this->flat();
                    xstr_sn_.expand(group);
                    std::vector<double> sn;
                    for (int ip = 0; ip < fp.n_plane; ip++)
                        for (int ic = 0; ic < fp.n_cell_plane; ic++)
                            sn.push_back(xstr_sn_[ic + fp.plane_xs_offset[ip]]);
                    out.put(p + "sn_xs", sn);
                    const int n_ang2 = ang_quad_.ndir() / 2, n_cell = corrections_->n_cell();
                    std::vector<double> al((size_t)n_ang2 * n_cell * 2), be((size_t)n_ang2 * n_cell);
                    for (int ia = 0; ia < n_ang2; ia++)
                        for (int ic = 0; ic < n_cell; ic++) {
                            al[((size_t)ia * n_cell + ic) * 2 + 0] = corrections_->alpha(ic, ia, group, Normal::X_NORM);
                            al[((size_t)ia * n_cell + ic) * 2 + 1] = corrections_->alpha(ic, ia, group, Normal::Y_NORM);
                            be[(size_t)ia * n_cell + ic]           = corrections_->beta(ic, ia, group);
                        }
                    out.put(p + "alpha", al);
                    out.put(p + "beta", be);
                }
                n_rec++;
            }
        }
    }
};


std::set<RecordKey> parse_records(const std::string &s);

template <class SW> int run_golden(SW &sw, Source &source_ref, const std::string &out_path, int outers,
                                   const std::string &records)
{
    Source *source = &source_ref;
    const int ng   = sw.n_group();
    auto fp        = sw.flat();
    fp.to_arrayfile().save(out_path + ".mocflat");
    ArrayFile out;
    out.put_scalar<int32_t>("n_inner", sw.n_inner());
    out.put_scalar<int32_t>("gs_boundary", sw.gs() ? 1 : 0);
    out.put_scalar<int32_t>("n_outer", outers);
    for (int g = 0; g < ng; g++) {
        out.put("xs_tr_" + std::to_string(g), sw.xs_per_reg(g, 0));
        out.put("xs_self_" + std::to_string(g), sw.xs_per_reg(g, 1));
        out.put("xs_nf_" + std::to_string(g), sw.xs_per_reg(g, 2));
        out.put("xs_ch_" + std::to_string(g), sw.xs_per_reg(g, 3));
        out.put("xs_scat_to_" + std::to_string(g), sw.xs_scat_to(g));
    }
    std::set<RecordKey> want = parse_records(records);
    sw.initialize();
    real_t k = 1.0;
    ArrayB1 fs(sw.n_reg());
    int n_rec = 0;
    std::vector<double> khist;
    for (int outer = 0; outer < outers; outer++) {
        sw.calc_fission_source(k, fs);
        { // inputs and result of TransportSweeper::calc_fission_source (transport_sweeper.cpp:119-134) for this outer
            const ArrayB2 &fl = sw.flux();
            std::vector<double> f0(fl.begin(), fl.end()), fsv(fs.begin(), fs.end());
            const std::string o = "outer" + std::to_string(outer) + "_";
            out.put(o + "flux_start", f0.data(), {(uint64_t)fl.extent(0), (uint64_t)fl.extent(1)});
            out.put(o + "fission_source", fsv);
            out.put_scalar<double>(o + "k", (double)k);
        }
        sw.store_old_flux();
        for (int ig = 0; ig < ng; ig++) {
            source->initialize_group(ig);
            source->fission(fs, ig);
            source->in_scatter(ig);
            sw.sweep_recorded(ig, outer, want, out, n_rec);
        }
        k = k * sw.total_fission(false) / sw.total_fission(true);
        khist.push_back(k);
    }
    out.put_scalar<int32_t>("n_rec", n_rec);
    out.put("k_history", khist);
    const ArrayB2 &flux = sw.flux();
    std::vector<double> f(flux.begin(), flux.end());
    out.put("flux_final", f.data(), {(uint64_t)flux.extent(0), (uint64_t)flux.extent(1)});
    out.save(out_path + ".golden");
    std::printf("golden: %d records, k after %d outers = %.12f\n", n_rec, outers, (double)k);
    return 0;
}

void set_xml(pugi::xml_document &doc, const std::string &spec)
{
    // path/to/node@attr=value
    auto at = spec.find('@');
    auto eq = spec.find('=', at);
    if (at == std::string::npos || eq == std::string::npos)
        throw std::runtime_error("bad --set spec: " + spec);
    std::string path = spec.substr(0, at), attr = spec.substr(at + 1, eq - at - 1),
                val = spec.substr(eq + 1);
    pugi::xml_node node = doc;
    std::stringstream ss(path);
    std::string part;
    while (std::getline(ss, part, '/')) {
        pugi::xml_node c = node.child(part.c_str());
        if (c.empty())
            c = node.append_child(part.c_str());
        node = c;
    }
    if (node.attribute(attr.c_str()).empty())
        node.append_attribute(attr.c_str());
    node.attribute(attr.c_str()).set_value(val.c_str());
}

std::set<RecordKey> parse_records(const std::string &s)
{
    std::set<RecordKey> out;
    std::stringstream ss(s);
    std::string item;
    while (std::getline(ss, item, ',')) {
        RecordKey k;
        if (std::sscanf(item.c_str(), "%d:%d:%d", &k.outer, &k.group, &k.inner) != 3)
            throw std::runtime_error("bad --records item: " + item);
        out.insert(k);
    }
    return out;
}

double total_fission(const TransportSweeper &sw, bool old)
{
    return sw.total_fission(old);
}
}

int main(int argc, char **argv)
{
    try {
        if (argc < 3) {
            std::cerr << "usage: ref_tool flat|golden|solve|time <in.xml> [out] [options]\n";
            return 2;
        }
        std::string cmd = argv[1], xml = argv[2];
        std::string out_path;
        std::vector<std::string> sets;
        int outers = 1, sweeps = 2, warmup = 1, groups_limit = 0;
        bool cmfd = false, twod3d = false;
        std::string records;
        for (int i = 3; i < argc; i++) {
            std::string a = argv[i];
            if (a == "--set" && i + 1 < argc)
                sets.push_back(argv[++i]);
            else if (a == "--outers" && i + 1 < argc)
                outers = std::atoi(argv[++i]);
            else if (a == "--sweeps" && i + 1 < argc)
                sweeps = std::atoi(argv[++i]);
            else if (a == "--warmup" && i + 1 < argc)
                warmup = std::atoi(argv[++i]);
            else if (a == "--records" && i + 1 < argc)
                records = argv[++i];
            else if (a == "--groups" && i + 1 < argc)
                groups_limit = std::atoi(argv[++i]);
            else if (a == "--cmfd")
                cmfd = true;
            else if (a == "--2d3d")
                twod3d = cmfd = true;
            else if (out_path.empty())
                out_path = a;
            else
                throw std::runtime_error("unknown argument: " + a);
        }

        pugi::xml_document doc;
        auto res = doc.load_file(xml.c_str());
        if (!res)
            throw std::runtime_error("cannot parse " + xml + ": " + res.description());
        for (const auto &s : sets)
            set_xml(doc, s);

        StartLogFile("ref_tool");
        RootTimer.tic();

        if (cmd == "solve") {
            CoreMesh mesh(doc);
            auto solver = SolverFactory(doc.child("solver"), mesh);
            auto t0     = std::chrono::steady_clock::now();
            solver->solve();
            double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
            const TransportSweeper *sw = solver->sweeper();
            ArrayFile out;
            const ArrayB2 &flux = sw->flux();
            std::vector<double> f(flux.begin(), flux.end());
            out.put("flux", f.data(), {(uint64_t)flux.extent(0), (uint64_t)flux.extent(1)});
            // k is not exposed by the Solver interface: recompute the final update ratio is not
            // possible either, so parse it from the convergence log the reference prints. The
            // caller reads k from stdout; here we only store flux and timing.
            out.put_scalar<double>("solve_seconds", secs);
            ArrayB3 pp = sw->pin_powers();
            std::vector<double> ppv(pp.begin(), pp.end());
            out.put("pin_powers", ppv);
            if (!out_path.empty())
                out.save(out_path);
            RootTimer.toc();
            std::cout << RootTimer << std::endl;
            return 0;
        }

        CoreMesh mesh(doc);
        pugi::xml_node solver_node = doc.child("solver");
        std::unique_ptr<ExposedMoC> sw_moc;
        std::unique_ptr<Exposed2D3D> sw_2d3d;
        if (twod3d)
            sw_2d3d.reset(new Exposed2D3D(solver_node.child("sweeper"), mesh));
        else
            sw_moc.reset(new ExposedMoC(solver_node.child("sweeper"), mesh));
        if (twod3d && cmd != "golden")
            throw std::runtime_error("--2d3d only applies to the golden command");
        if (twod3d) {
            Exposed2D3D &sw = *sw_2d3d;
            UP_Source_t source = sw.create_source(solver_node.child("source"));
            sw.assign_source(source.get());
            std::unique_ptr<CoarseData> cd(new CoarseData(mesh, sw.n_group()));
            sw.set_coarse_data(cd.get());
            return run_golden(sw, *source, out_path, outers, records);
        }
        ExposedMoC &sw = *sw_moc;

        if (cmd == "flat") {
            sw.flat().to_arrayfile().save(out_path);
            return 0;
        }

        UP_Source_t source = sw.create_source(solver_node.child("source"));
        sw.assign_source(source.get());
        std::unique_ptr<CoarseData> cd;
        if (cmfd) {
            cd.reset(new CoarseData(mesh, sw.n_group()));
            sw.set_coarse_data(cd.get());
        }
        const int ng = sw.n_group();

        if (cmd == "golden")
            return run_golden(sw, *source, out_path, outers, records);

        if (cmd == "time") {
            // Time `sweeps` full passes (every group once, n_inner inners each) of the
            // reference sweeper, after one untimed warm-up pass.
            auto fp         = sw.flat();
            int64_t S       = fp.n_seg_reference;
            sw.initialize();
            ArrayB1 fs(sw.n_reg());
            sw.calc_fission_source(1.0, fs);
            // --groups N: a bounded sample, the first N groups of every pass (all groups cost the same:
            // same rays, same n_inner, same share of tallying inners)
            const int ng_pass = (groups_limit > 0 && groups_limit < ng) ? groups_limit : ng;
            auto pass = [&]() {
                for (int ig = 0; ig < ng_pass; ig++) {
                    source->initialize_group(ig);
                    source->fission(fs, ig);
                    source->in_scatter(ig);
                    sw.sweep(ig);
                }
            };
            for (int i = 0; i < warmup; i++)
                pass();
            auto t0 = std::chrono::steady_clock::now();
            for (int i = 0; i < sweeps; i++)
                pass();
            double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
            double updates = 2.0 * (double)S * ng_pass * sw.n_inner() * sweeps;
            std::printf("{\"impl\": \"reference\", \"segments\": %lld, \"n_reg\": %d, \"groups\": %d, \"n_inner\": %d, "
                        "\"passes\": %d, \"updates\": %.0f, \"seconds\": %.6f, \"updates_per_s\": %.6e, "
                        "\"threads\": %d, \"current_tally\": %s}\n",
                        (long long)S, (int)sw.n_reg(), ng_pass, sw.n_inner(), sweeps, updates, secs, updates / secs,
                        omp_get_max_threads(), cmfd ? "true" : "false");
            return 0;
        }
        throw std::runtime_error("unknown command: " + cmd);
    } catch (const mocc::Exception &e) {
        std::cerr << "ref_tool: reference exception: " << e.what() << std::endl;
        return 1;
    } catch (const std::exception &e) {
        std::cerr << "ref_tool: " << e.what() << std::endl;
        return 1;
    }
}
