// ORACLE — test infrastructure only.  Stand-in for Assimp (absent from the image and from the reference tree):
// only the names src/core/Resource.h:29 declares against.  Model import (src/core/Resource.cpp) is NOT built;
// oracle/ref_host.cpp feeds mesh arrays to the reference's Scene through Resource::openModelInstance instead.
#pragma once
struct aiMesh; struct aiScene; struct aiNode;
