// slk_geometry.h -- the (team threads, CTA threads) geometries the peel kernels are compiled for.
// The plan builder picks one (slk_plan.cc), the launcher dispatches on it (slk_capi.cu).
#ifndef SLK_GEOMETRY_H
#define SLK_GEOMETRY_H

// X(team threads, max CTA threads)
#define SLK_GEOMETRIES(X) \
    X(32, 128) X(32, 384) X(64, 128) X(64, 384) X(64, 512) X(96, 576) X(128, 384) X(128, 640) X(192, 576) X(256, 768) X(512, 512)

// largest CTA a team size is compiled for (0: unsupported team size)
static inline int slk_max_cta(int team) {
    int best = 0;
#define SLK_X(T, C) if(team == T && C > best) best = C;
    SLK_GEOMETRIES(SLK_X)
#undef SLK_X
    return best;
}

#endif
