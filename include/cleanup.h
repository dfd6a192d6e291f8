// cleanup.h -- shutdown half of the reference-shaped C++ facade over libmedseg_b200.so.
//
// Interface kept: the one call src/main.cpp:185 makes at `exit` (declared at /root/reference/include/cleanup.h:7,
// implemented at src/cleanup.cpp:10-64).  Behaviour: destroys the process-wide ms_handle created by
// MedicalSeg::initialize_engine (streams, weights, activation buffers, stage workspaces), appends the reference's
// "=== Cleaning Up Resources ===" / "All resources cleaned up successfully" lines to the log, closes the log stream and
// prints "Resources cleaned up successfully".  Safe to call without a prior initialize_engine and safe to call twice
// (the reference double-destroys its engine; this one does not).
#pragma once

namespace MedicalSeg {
void cleanup_resources();
}
