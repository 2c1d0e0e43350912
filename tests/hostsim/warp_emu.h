// warp_emu.h — TEST-ONLY 32-lane warp emulator for the CPU suite.  Each lane is a fibre (own stack, hand-written
// x86-64 context switch); a lane runs until it reaches a warp collective (ballot / shfl / sync), the scheduler
// resumes the next one, and when every live lane has arrived the collective is evaluated and all continue.
// Lanes never run concurrently, so "shared memory" is ordinary memory; to expose missing warp syncs the lane
// order can be reversed (a lane that reads what a later-scheduled lane writes then sees stale data).
#pragma once
#include <stdint.h>

#include <functional>

namespace wemu {
// run body(lane) for lanes 0..31 as one warp; order: 0 = ascending lane schedule, 1 = descending
void run_warp(const std::function<void(uint32_t)>& body, int order);
uint64_t collectives();   // collectives evaluated by the last run_warp on this thread
}
