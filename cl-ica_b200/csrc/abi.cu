// abi.cu -- library-wide pieces of the C ABI: version, thread-local error string, device query.
#include "common.cuh"

#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <mutex>

namespace clica {

int env_flag(const char* name, int dflt) {
    const char* v = getenv(name);
    return v ? atoi(v) : dflt;
}

char* last_error_buffer() {
    static thread_local char buf[512] = {0};
    return buf;
}

int get_device_info(DeviceInfo* out) {
    static DeviceInfo cache[64];
    int dev = 0;
    CLICA_CUDA_OK(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64) return fail(CLICA_E_BADARG, "device ordinal %d out of range", dev);
    if (!cache[dev].ok) {
        DeviceInfo di;
        CLICA_CUDA_OK(cudaDeviceGetAttribute(&di.sm_count, cudaDevAttrMultiProcessorCount, dev));
        CLICA_CUDA_OK(cudaDeviceGetAttribute(&di.cc_major, cudaDevAttrComputeCapabilityMajor, dev));
        CLICA_CUDA_OK(cudaDeviceGetAttribute(&di.cc_minor, cudaDevAttrComputeCapabilityMinor, dev));
        if (di.cc_major != 10)
            return fail(CLICA_E_ARCH, "libclica_sm100 is built for sm_100a only; device %d is sm_%d%d", dev,
                        di.cc_major, di.cc_minor);
        di.ok = 1;
        cache[dev] = di;
    }
    *out = cache[dev];
    return 0;
}

namespace { std::atomic<int> g_sm_reserve{0}; }
int tc_sm_reserve() { return g_sm_reserve.load(std::memory_order_relaxed); }

// ---- launch accounting ---------------------------------------------------------------------------------
namespace {
constexpr int kProfSlots = 8192;
struct ProfState {
    std::atomic<long long> launches[kNumFamilies];
    std::atomic<int> enabled{0};
    std::atomic<int> next{0};
    cudaEvent_t start[kProfSlots];
    cudaEvent_t stop[kProfSlots];
    int family[kProfSlots];
    bool have_events = false;
    std::mutex mu;
};
ProfState& prof() {
    static ProfState st;
    return st;
}
}  // namespace

LaunchScope::LaunchScope(cudaStream_t st, int family, int launches) : st_(st), slot_(-1) {
    ProfState& ps = prof();
    ps.launches[family].fetch_add(launches, std::memory_order_relaxed);
    if (ps.enabled.load(std::memory_order_relaxed)) {
        int slot = ps.next.fetch_add(1);
        if (slot < kProfSlots) {
            slot_ = slot;
            ps.family[slot] = family;
            cudaEventRecord(ps.start[slot], st);
        }
    }
}
LaunchScope::~LaunchScope() {
    if (slot_ >= 0) cudaEventRecord(prof().stop[slot_], st_);
}

}  // namespace clica

extern "C" long long clica_launch_count(int family) {
    using namespace clica;
    long long n = 0;
    for (int f = 0; f < kNumFamilies; ++f)
        if (family < 0 || family == f) n += prof().launches[f].load();
    return n;
}

extern "C" int clica_prof_enable(int on) {
    using namespace clica;
    ProfState& ps = prof();
    std::lock_guard<std::mutex> lk(ps.mu);
    if (on && !ps.have_events) {
        for (int i = 0; i < kProfSlots; ++i) {
            CLICA_CUDA_OK(cudaEventCreate(&ps.start[i]));
            CLICA_CUDA_OK(cudaEventCreate(&ps.stop[i]));
        }
        ps.have_events = true;
    }
    ps.next.store(0);
    ps.enabled.store(on ? 1 : 0);
    return 0;
}

extern "C" int clica_prof_collect(float* ms_by_family, int* scopes_by_family) {
    using namespace clica;
    ProfState& ps = prof();
    std::lock_guard<std::mutex> lk(ps.mu);
    CLICA_REQUIRE(ms_by_family && scopes_by_family, CLICA_E_BADARG, "prof_collect: null pointer");
    for (int f = 0; f < kNumFamilies; ++f) { ms_by_family[f] = 0.f; scopes_by_family[f] = 0; }
    int n = ps.next.load();
    if (n > kProfSlots) n = kProfSlots;
    for (int i = 0; i < n; ++i) {
        CLICA_CUDA_OK(cudaEventSynchronize(ps.stop[i]));
        float ms = 0.f;
        CLICA_CUDA_OK(cudaEventElapsedTime(&ms, ps.start[i], ps.stop[i]));
        ms_by_family[ps.family[i]] += ms;
        scopes_by_family[ps.family[i]] += 1;
    }
    ps.next.store(0);
    return 0;
}

extern "C" int clica_tc_set_sm_reserve(int sms) {
    CLICA_REQUIRE(sms >= 0 && sms <= 64, CLICA_E_BADARG, "tc_set_sm_reserve: %d outside [0, 64]", sms);
    const int prev = clica::g_sm_reserve.exchange(sms);
    return prev < 0 ? 0 : 0;
}

extern "C" int clica_abi_version(void) { return CLICA_ABI_VERSION; }

extern "C" const char* clica_last_error(void) { return clica::last_error_buffer(); }

extern "C" int clica_device_info(int* sm_count, int* cc_major, int* cc_minor) {
    clica::DeviceInfo di;
    int rc = clica::get_device_info(&di);
    if (rc) return rc;
    if (sm_count) *sm_count = di.sm_count;
    if (cc_major) *cc_major = di.cc_major;
    if (cc_minor) *cc_minor = di.cc_minor;
    return 0;
}
