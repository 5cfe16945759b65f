// mem_pool.cpp -- see mem_pool.h.
#include "mem_pool.h"
#include <cstdlib>
#include <mutex>
#include <unordered_map>
#include <vector>

namespace goicp {
namespace {

struct Block { void* p; size_t bytes; int device; };     // device < 0: pinned host memory

std::mutex g_mtx;
std::unordered_map<void*, Block> g_live;                 // blocks handed out
std::vector<Block> g_free;                               // cached blocks
size_t g_cached_dev = 0, g_cached_host = 0;

size_t cap_bytes()
{
    static const size_t cap = [] {
        const char* e = std::getenv("GOICP_POOL_MAX_MB");
        const double mb = e ? std::atof(e) : 8192.0;
        return (size_t)(mb < 0 ? 0 : mb) << 20;
    }();
    return cap;
}
size_t round_up(size_t b) { return (b + 511) & ~(size_t)511; }

// best fit among cached blocks of the right kind: smallest block >= bytes that wastes at most half of itself
bool take_cached(size_t bytes, int device, Block* out)
{
    int best = -1;
    for (int i = 0; i < (int)g_free.size(); i++) {
        const Block& b = g_free[i];
        if (b.device != device || b.bytes < bytes || b.bytes > 2 * bytes + (1u << 20)) continue;
        if (best < 0 || b.bytes < g_free[best].bytes) best = i;
    }
    if (best < 0) return false;
    *out = g_free[best];
    g_free.erase(g_free.begin() + best);
    (device < 0 ? g_cached_host : g_cached_dev) -= out->bytes;
    return true;
}

} // namespace

cudaError_t pool_alloc(void** p, size_t bytes)
{
    *p = nullptr;
    if (bytes == 0) return cudaSuccess;
    bytes = round_up(bytes);
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    std::lock_guard<std::mutex> lk(g_mtx);
    Block b;
    if (!take_cached(bytes, dev, &b)) {
        e = cudaMalloc(&b.p, bytes);
        if (e != cudaSuccess) {                      // out of memory: drop the cache of this device and retry once
            (void)cudaGetLastError();
            for (size_t i = 0; i < g_free.size();) {
                if (g_free[i].device == dev) { cudaFree(g_free[i].p); g_cached_dev -= g_free[i].bytes; g_free.erase(g_free.begin() + i); }
                else i++;
            }
            e = cudaMalloc(&b.p, bytes);
            if (e != cudaSuccess) return e;
        }
        b.bytes = bytes; b.device = dev;
    }
    g_live[b.p] = b;
    *p = b.p;
    return cudaSuccess;
}

void pool_free(void* p)
{
    if (!p) return;
    std::lock_guard<std::mutex> lk(g_mtx);
    auto it = g_live.find(p);
    if (it == g_live.end()) { cudaFree(p); return; }
    const Block b = it->second;
    g_live.erase(it);
    if (g_cached_dev + b.bytes > cap_bytes()) {
        int cur = 0; cudaGetDevice(&cur);
        if (cur != b.device) cudaSetDevice(b.device);
        cudaFree(b.p);
        if (cur != b.device) cudaSetDevice(cur);
        return;
    }
    g_free.push_back(b); g_cached_dev += b.bytes;
}

cudaError_t pool_alloc_host(void** p, size_t bytes)
{
    *p = nullptr;
    if (bytes == 0) return cudaSuccess;
    bytes = round_up(bytes);
    std::lock_guard<std::mutex> lk(g_mtx);
    Block b;
    if (!take_cached(bytes, -1, &b)) {
        cudaError_t e = cudaMallocHost(&b.p, bytes);
        if (e != cudaSuccess) return e;
        b.bytes = bytes; b.device = -1;
    }
    g_live[b.p] = b;
    *p = b.p;
    return cudaSuccess;
}

void pool_free_host(void* p)
{
    if (!p) return;
    std::lock_guard<std::mutex> lk(g_mtx);
    auto it = g_live.find(p);
    if (it == g_live.end()) { cudaFreeHost(p); return; }
    const Block b = it->second;
    g_live.erase(it);
    if (g_cached_host + b.bytes > ((size_t)256 << 20)) { cudaFreeHost(b.p); return; }
    g_free.push_back(b); g_cached_host += b.bytes;
}

void pool_trim()
{
    std::lock_guard<std::mutex> lk(g_mtx);
    int cur = 0; cudaGetDevice(&cur);
    for (const Block& b : g_free) {
        if (b.device < 0) cudaFreeHost(b.p);
        else { cudaSetDevice(b.device); cudaFree(b.p); }
    }
    cudaSetDevice(cur);
    g_free.clear(); g_cached_dev = g_cached_host = 0;
}

} // namespace goicp
