// Device-resident state block of the fused optimizer (shared by optim.cu and optim_p2p.cu).
#pragma once

struct OptState {        // lives in device memory, 8 x 4 bytes
    float scale;         // current loss scale
    int found_inf;       // set by k_grads_check for the current step
    int growth_tracker;  // consecutive finite steps since the last scale change
    int good_steps;      // optimizer steps actually taken (bias correction / LambdaLR epoch)
    int pad[4];
};

