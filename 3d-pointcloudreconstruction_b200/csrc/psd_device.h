// psd_device.h -- per-device host state of libpsd_b200.so.
//
// Everything the launchers cache is keyed by the CUDA device that is current at the call (one process may drive
// several GPUs, and several host threads may call in): SM count / shared-memory limit, the one-shot
// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) opt-ins (an attribute applies per device), and the library-owned
// device workspaces.  All of it sits behind one mutex; the test / measurement switches are atomics.
#pragma once
#include <cuda_runtime.h>

#include <atomic>
#include <mutex>

namespace psd {

constexpr int kMaxDevices = 64;
constexpr int kStepSlots = 8;     // workspaces of the host-buffer training step (psd_chamfer_loss_step_host_ex)
constexpr int kMergeSlots = 16;   // (stream -> workspace) entries of the multi-tile NN merge workspace

struct MergeWorkspace {           // 64-bit (distance, index) keys of chamfer_nn_tc's multi-tile mode, always left all-ones
    cudaStream_t stream = nullptr;
    unsigned long long *ptr = nullptr;
    size_t elems = 0;
    bool used = false;
    bool dirty = false;          // keys may be left behind (fresh allocation, failed launch): refill before use
    unsigned long long last_use = 0;
};

struct DeviceState {
    bool init = false;
    int device = -1;
    int num_sms = 0, max_smem = 0;
    bool attr_nn = false, attr_tc = false, attr_proj = false;   // dynamic shared-memory opt-ins done on this device
    float *fwd_ws = nullptr;                                    // psd_chamfer_forward_host staging
    size_t fwd_ws_bytes = 0;
    float *step_ws[kStepSlots] = {};                            // psd_chamfer_loss_step_host_ex staging, by slot
    size_t step_ws_bytes[kStepSlots] = {};
    MergeWorkspace merge[kMergeSlots];
    unsigned long long merge_clock = 0;
};

// State of the device that is current on the calling thread (initialised on first use).  Returns nullptr and sets *err on a
// CUDA failure or a device ordinal >= kMaxDevices.  The caller holds state_mutex() while it reads or writes anything but
// num_sms / max_smem (those never change once `init` is set).
DeviceState *device_state(cudaError_t *err);
std::mutex &state_mutex();

}  // namespace psd
