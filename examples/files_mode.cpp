// files-mode driver in C++: what Reader::run (files mode, /root/reference/test/reader.cpp:449-505) + DProcessor::run
// (/root/reference/test/processor.cpp:156-174) do for one recorded frame, on top of the B200 front-end.
//   usage: files_mode <NNNN_cloud.pcd> [show_mode=10]
// Prints one line per hand-off plane: type count a b c d cx cy cz  (floats as hex so a test can compare bit for bit).
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <sstream>
#include <string>

#include "StairPerception_b200.hpp"

static bool load_pcd(const std::string &path, CloudType &cloud)
{
    std::ifstream f(path, std::ios::binary);
    if (!f) return false;
    std::string line;
    size_t points = 0;
    while (std::getline(f, line)) {
        std::istringstream is(line);
        std::string key;
        is >> key;
        if (key == "WIDTH") is >> cloud.width;
        else if (key == "HEIGHT") is >> cloud.height;
        else if (key == "POINTS") is >> points;
        else if (key == "FIELDS") { std::string a, b, c, d; is >> a >> b >> c >> d; if (a != "x" || b != "y" || c != "z" || d != "rgba") return false; }
        else if (key == "DATA") { std::string kind; is >> kind; if (kind != "binary") return false; break; }
    }
    struct Rec { float x, y, z; uint32_t rgba; };
    std::vector<Rec> rec(points);
    f.read(reinterpret_cast<char *>(rec.data()), (std::streamsize)(points * sizeof(Rec)));
    if ((size_t)f.gcount() != points * sizeof(Rec)) return false;
    cloud.points.resize(points);
    for (size_t i = 0; i < points; ++i) { PointType p = PointType(); p.x = rec[i].x; p.y = rec[i].y; p.z = rec[i].z; p.rgba = rec[i].rgba; cloud.points[i] = p; }
    cloud.is_dense = false;
    return true;
}

int main(int argc, char **argv)
{
    if (argc < 2) { std::fprintf(stderr, "usage: %s <cloud.pcd> [show_mode]\n", argv[0]); return 2; }
    const ShowModeDef mode = argc > 2 ? (ShowModeDef)std::atoi(argv[2]) : stair;
    CloudType cloud_in, cloud_show;
    pcl::PointCloud<pcl::Normal> normal_show;
    if (!load_pcd(argv[1], cloud_in)) { std::fprintf(stderr, "cannot read %s\n", argv[1]); return 2; }
    std::vector<Plane> vector_plane_sorted;
    try {
        StairDetection stair_detection(sp_b200::default_parameters());          // a new object per frame, as test/processor.cpp:156
        sp_stair stair_model;
        const bool has_stair = stair_detection.process(cloud_in, mode, vector_plane_sorted, cloud_show, normal_show, nullptr, &stair_model);
        const bool has = mode == stair ? !vector_plane_sorted.empty() : false;
        std::printf("has_planes %d planes %zu cloud_show %zu\n", has ? 1 : 0, vector_plane_sorted.size(), cloud_show.points.size());
        for (const Plane &p : vector_plane_sorted) {
            std::printf("%d %zu", (int)p.type, p.cloud.points.size());
            for (size_t k = 0; k < p.coefficients.values.size(); ++k) std::printf(" %a", (double)p.coefficients.values[k]);
            std::printf(" %a %a %a\n", (double)p.center.x, (double)p.center.y, (double)p.center.z);
        }
        if (mode == stair) {
            // StairDetection::printStairEstParam, StairPerception.cpp:265-290
            std::printf("stair %d steps %d\nHeight\t\tDepth\t\tV_Distance\tH_Distance\n", has_stair ? 1 : 0, stair_model.n_steps);
            for (int i = 0; i < stair_model.n_nodes; ++i)
                if (stair_model.nodes[i].kind == 1)
                    std::printf("%g\t%g\t%g\t%g\n", stair_model.nodes[i].height, stair_model.nodes[i].depth, (double)stair_model.nodes[i].line_h, (double)stair_model.nodes[i].line_d);
            if (has_stair) {
                // what the GUI draws per plane (computePlaneCounter, StairPerception.cpp:2861-3075; gui.cpp:326-330): the closed contours
                std::printf("side planes: left %g %g %g %g  right %g %g %g %g\n", (double)stair_model.slide_left[0], (double)stair_model.slide_left[1],
                            (double)stair_model.slide_left[2], (double)stair_model.slide_left[3], (double)stair_model.slide_right[0], (double)stair_model.slide_right[1],
                            (double)stair_model.slide_right[2], (double)stair_model.slide_right[3]);
                for (size_t k = 0; k < vector_plane_sorted.size() && k < (size_t)SP_MAX_PLANES; ++k) {
                    if (stair_model.contour_width[k] != 4) continue;
                    std::printf("contour plane %zu:", k);
                    for (int c = 0; c < 4; ++c) std::printf(" (%.3f %.3f %.3f)", (double)stair_model.contour[k][c][0], (double)stair_model.contour[k][c][1], (double)stair_model.contour[k][c][2]);
                    std::printf("\n");
                }
            }
        }
    } catch (const std::exception &e) {
        std::fprintf(stderr, "error: %s\n", e.what());
        return 3;
    }
    return 0;
}
