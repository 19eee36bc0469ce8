/*
 * TEST INFRASTRUCTURE ONLY -- not part of the product.
 *
 * A headless stand-in for <SDL.h> so that the reference's src/Main.cpp can be
 * compiled where it lies (it only touches SDL for the window, the worker
 * threads and the mouse; the render path -- shade / renderTile / renderBatch,
 * Main.cpp:81-202 -- needs nothing but `SDL_Surface::pixels` and `::pitch`).
 * Written from scratch for this repository: just enough declarations for the
 * identifiers Main.cpp mentions. Nothing here is ever called on the render
 * path; `reference_main` is never invoked by the harness.
 */
#ifndef SVO_REF_SHIM_SDL_H_
#define SVO_REF_SHIM_SDL_H_

#include <stdlib.h>

struct SDL_Surface {
    void *pixels;
    int   pitch;
    int   w, h;
};
struct SDL_Thread { int unused; };

enum { SDL_INIT_VIDEO = 0x20, SDL_SWSURFACE = 0 };
enum { SDL_MOUSEMOTION = 4 };
enum { SDLK_ESCAPE = 27 };

#define SDL_MUSTLOCK(s) 0

static inline int  SDL_Init(unsigned) { return 0; }
static inline void SDL_Quit() {}
static inline void SDL_WM_SetCaption(const char *, const char *) {}
static inline int  SDL_LockSurface(SDL_Surface *) { return 0; }
static inline void SDL_UnlockSurface(SDL_Surface *) {}
static inline void SDL_UpdateRect(SDL_Surface *, int, int, unsigned, unsigned) {}
static inline SDL_Surface *SDL_SetVideoMode(int w, int h, int, unsigned) {
    SDL_Surface *s = (SDL_Surface *)malloc(sizeof(SDL_Surface));
    s->pixels = calloc((size_t)w*(size_t)h, 4);
    s->pitch = w*4;
    s->w = w;
    s->h = h;
    return s;
}
static inline SDL_Thread *SDL_CreateThread(int (*)(void *), void *) { return 0; }
static inline void SDL_WaitThread(SDL_Thread *, int *) {}

#endif
