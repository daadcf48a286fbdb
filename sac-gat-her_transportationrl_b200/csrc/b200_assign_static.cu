// Network-specialised kernel, built-in topology (Sioux Falls) + the registry of on-demand plugins.
// Kernel template: b200_assign_static.cuh; topology data: b200_topo_siouxfalls.h.
#include "b200_assign_internal.cuh"
#include "b200_topo_siouxfalls.h"

#include <dlfcn.h>
#include <mutex>
#include <string>

using namespace b200i;

namespace {
#include "b200_assign_static.cuh"

struct Plugin {
    std::string path;
    b200_plugin_prepare_fn prepare;
    b200_plugin_launch_fn launch;
};
std::vector<Plugin> g_plugins;
std::mutex g_plugins_mutex;
}  // namespace

int b200_launch_static(const b200_graph *g, const Params &p, int prec, cudaStream_t st)
{
    if (g->plugin_launch) return ((b200_plugin_launch_fn)g->plugin_launch)(g, &p, prec, st);
    return launch_static<TopoSiouxFalls>(g, p, prec, st);
}

// Eligibility (link list equals the built-in topology or a registered plugin's) + tables.
int b200_static_prepare(b200_graph *g, const int32_t *row, const int32_t *col, const float *t0, const float *cap0,
                        int num_od, const int32_t *od_o, const int32_t *od_d, const double *od_q,
                        const int32_t *origin_slot)
{
    g->static_ok = 0;
    g->plugin_launch = nullptr;
    if (static_try_prepare<TopoSiouxFalls>(g, row, col, t0, cap0, num_od, od_o, od_d, od_q, origin_slot)) return 1;
    std::lock_guard<std::mutex> lock(g_plugins_mutex);
    for (const Plugin &pl : g_plugins)
        if (pl.prepare(g, row, col, t0, cap0, num_od, od_o, od_d, od_q, origin_slot)) {
            g->plugin_launch = (void *)pl.launch;
            return 1;
        }
    return 0;
}

// A plugin is b200_static_plugin.cu compiled for one topology header (see _native.specialize).
extern "C" int b200_register_static_plugin(const char *path)
{
    B200_REQUIRE(path != nullptr, "null plugin path");
    std::lock_guard<std::mutex> lock(g_plugins_mutex);
    for (const Plugin &pl : g_plugins)
        if (pl.path == path) return 0;
    void *h = dlopen(path, RTLD_NOW | RTLD_LOCAL);
    if (!h) { b200_set_error("dlopen(%s): %s", path, dlerror()); return B200_EINVAL; }
    Plugin pl;
    pl.path = path;
    pl.prepare = (b200_plugin_prepare_fn)dlsym(h, "b200_plugin_prepare");
    pl.launch = (b200_plugin_launch_fn)dlsym(h, "b200_plugin_launch");
    int (*abi)(void) = (int (*)(void))dlsym(h, "b200_plugin_abi");
    if (!pl.prepare || !pl.launch || !abi || abi() != B200_PLUGIN_ABI) {
        b200_set_error("%s is not a static-kernel plugin of this library version", path);
        dlclose(h);
        return B200_EINVAL;
    }
    g_plugins.push_back(pl);
    return 0;
}
