// xs_updater.hpp -- XSMeshHomogenized::update() with the pin loop spread over the host threads
// (xs_update_parallel.cpp). The tables flattened from the mesh belong to the updater, i.e. to the sweeper that
// holds it: no process-wide state.
#pragma once
#include <memory>

namespace mocc {
class XSMeshHomogenized;
}

namespace mocc_b200 {

struct XsUpdaterImpl;

class XsUpdater {
public:
    XsUpdater();
    ~XsUpdater();
    XsUpdater(const XsUpdater &)            = delete;
    XsUpdater &operator=(const XsUpdater &) = delete;
    // = xs.update() (xs_mesh_homogenized.cpp:176-197), bit-identical
    void update(mocc::XSMeshHomogenized &xs);

private:
    std::unique_ptr<XsUpdaterImpl> impl_;
};

} // namespace mocc_b200
