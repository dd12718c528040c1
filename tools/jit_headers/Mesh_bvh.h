// hand-written stand-in for Mesh.json once it goes through the scene BVH (1 mesh + 1 plane + 1 point light)
#define MRT_JIT_BVH 1
#define MRT_JIT_BOXPAIRS(X, XS, X1, CB, CE)
#define MRT_JIT_SPHERES(X)
#define MRT_JIT_PLANES(X)
#define MRT_JIT_BXFS(X)
#define MRT_JIT_MESHES(X)
#define MRT_JIT_MINBLOCKS 8
#define MRT_JIT_EMIT_BINARY 1
#define MRT_JIT_SKY_BLACK 1
#define MRT_JIT_REFINE_SPHERES 0
#define MRT_JIT_ROT 0
#define MRT_JIT_N_BOX 0
#define MRT_JIT_N_SPHERE 0
#define MRT_JIT_N_ABOX 0
#define MRT_JIT_N_BXF 0
#define MRT_JIT_N_MESH 1
#define MRT_JIT_N_LIGHTS 1
#define MRT_JIT_N_PLANE 1
#define MRT_JIT_FIRST_SPHERE 0
#define MRT_JIT_FIRST_PLANE 0
#define MRT_JIT_FIRST_BXF 1
#define MRT_JIT_FIRST_MESH 1
#define MRT_JIT_F 11u
