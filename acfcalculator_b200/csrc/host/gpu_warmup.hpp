// Starts the CUDA context creation (acfb200_warmup, 1-2 s on a multi-GPU box) on a helper thread so that it overlaps the
// parsing of the input file; joined before the first compute call and on every other way out of main().
#pragma once
#include <thread>

#include "../../../include/acf_b200.h"

namespace acfhost {

class GpuWarmup {
public:
    GpuWarmup() : th_([] { acfb200_warmup(); }) {}
    ~GpuWarmup() { wait(); }
    void wait()
    {
        if (th_.joinable()) th_.join();
    }

private:
    std::thread th_;
};

}  // namespace acfhost
