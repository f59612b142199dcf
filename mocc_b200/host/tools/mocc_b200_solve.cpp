// mocc_b200_solve -- run a MOCC input through the UNMODIFIED solver stack (SolverFactory ->
// EigenSolver / FixedSourceSolver, CMFD) with whatever sweeper the XML selects ("moc" = the
// reference CPU sweeper, "moc_cuda" = the B200 sweeper) and dump what parity checks and
// the bench need: k history, scalar flux, pin powers, sweep timers.
//
//   mocc_b200_solve <in.xml> <out.arrays> [--set path/to/node@attr=value]...
//
// The reference driver (src/driver.cpp) writes these through HDF5, which this image lacks.
#include <chrono>
#include <cstdio>
#include <iostream>
#include <string>
#include <vector>

#include "pugixml.hpp"

// EigenSolver keeps its convergence history private and only exports it through HDF5
// (eigen_solver.cpp:293-322); this tool reads it directly.
#define private public
#include "solvers/eigen_solver.hpp"
#undef private

#include "core/core_mesh.hpp"
#include "solvers/solver_factory.hpp"
#include "util/error.hpp"
#include "util/files.hpp"
#include "util/timers.hpp"

#include "arrayfile.hpp"
#include "cuda_moc_sweeper.hpp"
#include "xml_amend.hpp"

using namespace mocc;

int main(int argc, char **argv)
{
    try {
        if (argc < 3) {
            std::cerr << "usage: mocc_b200_solve <in.xml> <out.arrays> [--set path@attr=value]...\n";
            return 2;
        }
        std::vector<std::string> sets;
        for (int i = 3; i < argc; i++) {
            const std::string a = argv[i];
            if (a == "--set" && i + 1 < argc)
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

        StartLogFile("mocc_b200_solve");
        RootTimer.tic();
        const auto t_begin = std::chrono::steady_clock::now();
        CoreMesh mesh(doc);
        SP_Solver_t solver = SolverFactory(doc.child("solver"), mesh);
        const auto t_setup = std::chrono::steady_clock::now();
        solver->solve();
        const auto t_end = std::chrono::steady_clock::now();
        RootTimer.toc();

        mocc_b200::ArrayFile out;
        const TransportSweeper *sw = solver->sweeper();
        const ArrayB2 &flux        = sw->flux();
        std::vector<double> f(flux.begin(), flux.end());
        out.put("flux", f.data(), {(uint64_t)flux.extent(0), (uint64_t)flux.extent(1)});
        ArrayB3 pp = sw->pin_powers();
        std::vector<double> ppv(pp.begin(), pp.end());
        out.put("pin_powers", ppv);
        std::vector<double> k, ek, epsi;
        if (const auto *es = dynamic_cast<const EigenSolver *>(solver.get())) {
            for (const auto &c : es->convergence_) {
                k.push_back(c.k), ek.push_back(c.error_k), epsi.push_back(c.error_psi);
            }
        }
        out.put("k_history", k);
        out.put("error_k", ek);
        out.put("error_psi", epsi);
        double sweep_s = 0.0;
        try {
            sweep_s = RootTimer["MoC Sweeper"]["Sweep"].time();
        } catch (...) {
        }
        out.put_scalar<double>("sweep_seconds", sweep_s);
        out.put_scalar<double>("setup_seconds", std::chrono::duration<double>(t_setup - t_begin).count());
        out.put_scalar<double>("solve_seconds", std::chrono::duration<double>(t_end - t_setup).count());
        double dev_ms = 0.0;
        if (const auto *cs = dynamic_cast<const mocc_b200::CudaMoCSweeper *>(sw))
            dev_ms = cs->device_sweep_ms();
        out.put_scalar<double>("device_sweep_ms", dev_ms);
        out.save(argv[2]);
        std::cout << RootTimer << std::endl;
        std::printf("mocc_b200_solve: outers=%zu k=%.12f sweep_seconds=%.4f device_sweep_ms=%.3f solve_seconds=%.3f\n",
                    k.size(), k.empty() ? 0.0 : k.back(), sweep_s, dev_ms,
                    std::chrono::duration<double>(t_end - t_setup).count());
        return 0;
    } catch (const mocc::Exception &e) {
        std::cerr << "mocc_b200_solve: " << e.what() << std::endl;
        return 1;
    } catch (const std::exception &e) {
        std::cerr << "mocc_b200_solve: " << e.what() << std::endl;
        return 1;
    }
}
