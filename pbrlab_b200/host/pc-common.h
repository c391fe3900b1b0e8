// Scene assembly of the command-line front-end (reference pc/pc-common.h:99, pc/pc-common.cc:100-270):
// argv[1..] = .obj / .hair files -> Scene (one local scene + one identity instance per shape), committed.
#ifndef PBRLAB_B200_PC_COMMON_H_
#define PBRLAB_B200_PC_COMMON_H_
#include "scene.h"
// commit_to_device = false stops after the host half of CommitScene (used by the GPU-less tests).
bool CreateScene(int argc, char** argv, pbrlab::Scene* scene, bool commit_to_device = true);
#endif  // PBRLAB_B200_PC_COMMON_H_
