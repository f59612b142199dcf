// flatten_ray_check -- TEST INFRASTRUCTURE ONLY.
// Traces the rays of the reference's own ray tests (src/sweepers/moc/tests/test_Ray.cpp:45-91 "simple_ray" on
// tests/6x5.xml, :145-305 "weird_ray" on one 10 cm pin cut into 80 x 80 regions) with the UNMODIFIED reference
// (moc::Ray, CoreMesh), pushes every ray through the flattener the B200 sweeper uses
// (mocc_b200::append_ray, mocc_b200/host/flatten.cpp) and prints the flat arrays as JSON, so that
// tests/test_flatten_rays.py can hold them against the known answers of test_Ray.cpp.
//   usage (cwd must hold c5g7.xsl): flatten_ray_check <reference tests dir> <out.json>
#include <cstdio>
#include <iostream>
#include <sstream>
#include <string>

#include "pugixml.hpp"

#include "core/core_mesh.hpp"
#include "core/geometry/points.hpp"
#include "ray.hpp"

#include "flatten.hpp"

using namespace mocc;
using mocc::moc::Ray;

static FILE *g_out = nullptr;
#define printf(...) fprintf(g_out, __VA_ARGS__)

static void dump(const char *name, const mocc_b200::FlatProblem &fp, bool last)
{
    printf("  \"%s\": {\n", name);
    auto ivec = [](const char *k, const auto &v, bool comma) {
        printf("    \"%s\": [", k);
        for (size_t i = 0; i < v.size(); i++)
            printf("%s%lld", i ? ", " : "", (long long)v[i]);
        printf("]%s\n", comma ? "," : "");
    };
    ivec("trk_bc", fp.trk_bc, true);
    ivec("trk_cm_start", fp.trk_cm_start, true); // cm_cell_fw, cm_cell_bw, cm_surf_fw, cm_surf_bw
    ivec("trk_seg_begin", fp.trk_seg_begin, true);
    ivec("trk_cm_begin", fp.trk_cm_begin, true);
    ivec("seg_fsr", fp.seg_fsr, true);
    ivec("cm_data", fp.cm_data, true);
    printf("    \"seg_len\": [");
    for (size_t i = 0; i < fp.seg_len.size(); i++)
        printf("%s%.17g", i ? ", " : "", fp.seg_len[i]);
    printf("]\n  }%s\n", last ? "" : ",");
}

int main(int argc, char **argv)
{
    if (argc < 3 || !(g_out = std::fopen(argv[2], "w"))) {
        std::cerr << "usage: flatten_ray_check <reference tests dir> <out.json>\n";
        return 2;
    }
    try {
        printf("{\n");
        {
            pugi::xml_document xml;
            if (!xml.load_file((std::string(argv[1]) + "/6x5.xml").c_str()))
                throw std::runtime_error("cannot read 6x5.xml");
            CoreMesh mesh(xml);
            mocc_b200::FlatProblem fp;
            // the two rays test_Ray.cpp:58-91 checks in detail
            mocc_b200::append_ray(fp, Ray(Point2(0.0, 1.0), Point2(4.0, 5.0), {{0, 0}}, 0, mesh));
            mocc_b200::append_ray(fp, Ray(Point2(4.0, 0.0), Point2(6.0, 2.0), {{0, 0}}, 0, mesh));
            dump("simple_ray", fp, false);
        }
        {
            // test_Ray.cpp:147-262: one pin of pitch 10 cm, rectangular mesh 80 x 80, one material
            std::stringstream x;
            x << "<mesh id=\"1\" type=\"rect\" pitch=\"10\"><sub_x>80</sub_x><sub_y>80</sub_y></mesh>\n<pin id=\"1\" mesh=\"1\">\n";
            for (int j = 0; j < 80; j++) {
                for (int i = 0; i < 80; i++)
                    x << " 1";
                x << "\n";
            }
            x << "</pin>\n<lattice id=\"1\" nx=\"1\" ny=\"1\">1</lattice>\n"
              << "<assembly id=\"1\" np=\"1\" hz=\"1.0\"><lattices>1</lattices></assembly>\n"
              << "<core nx=\"1\" ny=\"1\" north=\"reflect\" south=\"reflect\" top=\"reflect\" bottom=\"reflect\" "
                 "west=\"prescribed\" east=\"prescribed\">1</core>\n"
              << "<material_lib path=\"c5g7.xsl\"><material id=\"1\" name=\"UO2-3.3\" /></material_lib>\n";
            pugi::xml_document xml;
            if (!xml.load_string(x.str().c_str()))
                throw std::runtime_error("cannot parse the weird_ray mesh");
            CoreMesh mesh(xml);
            mocc_b200::FlatProblem fp;
            // test_Ray.cpp:270-277
            Point2 p1(-1.249999999999996 + 5.0, -5.0 + 5.0);
            Point2 p2(-5 + 5.0, -4.2424242424242422 + 5.0);
            mocc_b200::append_ray(fp, Ray(p1, p2, {{106, 7}}, 0, mesh));
            dump("weird_ray", fp, true);
        }
        printf("}\n");
        std::fclose(g_out);
    } catch (const std::exception &e) {
        std::cerr << "flatten_ray_check: " << e.what() << std::endl;
        return 1;
    }
    return 0;
}
