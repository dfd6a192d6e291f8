// cleanup.h -- drop-in for /root/reference/include/cleanup.h (line 7).
#ifndef CLEANUP_H
#define CLEANUP_H

namespace MedicalSeg {

// replaces cleanup_resources  (src/cleanup.cpp:10-64)
void cleanup_resources();

}  // namespace MedicalSeg

#endif  // CLEANUP_H
