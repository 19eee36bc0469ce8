/*
 * TEST INFRASTRUCTURE ONLY -- not part of the product.
 *
 * A headless stand-in for <SDL.h> so that the reference's src/Main.cpp can be
 * compiled where it lies (it only touches SDL for the window, the worker
 * threads and the mouse; the render path -- shade / renderTile / renderBatch,
 * Main.cpp:81-202 -- needs nothing but `SDL_Surface::pixels` and `::pitch`).
 * Written from scratch for this repository: just enough declarations for the
 * identifiers Main.cpp, Events.cpp and ThreadBarrier.cpp mention. Nothing here
 * is ever called on the render path.
 *
 * For the interactive path (row f4) the shim is a small working SDL: threads,
 * mutexes and semaphores on pthreads (so the reference's own ThreadBarrier and
 * its `-viewer` main loop run as written), a window that is a malloc'ed
 * surface, and an event queue that the harness fills from a script
 * (SDL_WaitEvent / SDL_PollEvent) -- every SDL_UpdateRect hands the finished
 * frame to the harness (oracle/ref_harness.cpp, svoref_viewer_run).
 */
#ifndef SVO_REF_SHIM_SDL_H_
#define SVO_REF_SHIM_SDL_H_

#include <stdlib.h>

struct SDL_Surface {
    void *pixels;
    int   pitch;
    int   w, h;
};
#include <pthread.h>
#include <semaphore.h>

enum { SDL_INIT_VIDEO = 0x20, SDL_SWSURFACE = 0 };

/* ---- events (the members src/Events.cpp reads) ---- */
enum { SDL_KEYDOWN = 2, SDL_KEYUP = 3, SDL_MOUSEMOTION = 4, SDL_MOUSEBUTTONDOWN = 5, SDL_MOUSEBUTTONUP = 6, SDL_QUIT = 12 };
enum { SDL_BUTTON_LEFT = 1, SDL_BUTTON_MIDDLE = 2, SDL_BUTTON_RIGHT = 3, SDL_BUTTON_WHEELUP = 4, SDL_BUTTON_WHEELDOWN = 5 };
enum { SDLK_ESCAPE = 27, SDLK_LAST = 323 };

struct SDL_MouseMotionEvent { unsigned char type; int x, y, xrel, yrel; };
struct SDL_MouseButtonEvent { unsigned char type; unsigned char button; };
struct SDL_keysym { int sym; };
struct SDL_KeyboardEvent { unsigned char type; SDL_keysym keysym; };
union SDL_Event {
    unsigned char type;
    SDL_MouseMotionEvent motion;
    SDL_MouseButtonEvent button;
    SDL_KeyboardEvent key;
};

/* implemented by the harness: the scripted queue and the frame sink */
extern "C" int  svo_shim_wait_event(SDL_Event *event);
extern "C" int  svo_shim_poll_event(SDL_Event *event);
extern "C" void svo_shim_present(SDL_Surface *surface);
extern "C" void svo_shim_surface_created(SDL_Surface *surface);

static inline int SDL_WaitEvent(SDL_Event *e) { return svo_shim_wait_event(e); }
static inline int SDL_PollEvent(SDL_Event *e) { return svo_shim_poll_event(e); }

/* ---- threads, mutexes, semaphores (src/ThreadBarrier.cpp, Main.cpp:346-371) ---- */
struct SDL_Thread { pthread_t handle; int (*fn)(void *); void *data; };
struct SDL_mutex { pthread_mutex_t m; };
struct SDL_sem { sem_t s; };

static inline void *svo_shim_thread_entry(void *p) {
    SDL_Thread *t = (SDL_Thread *)p;
    t->fn(t->data);
    return 0;
}
static inline SDL_Thread *SDL_CreateThread(int (*fn)(void *), void *data) {
    SDL_Thread *t = (SDL_Thread *)malloc(sizeof(SDL_Thread));
    t->fn = fn;
    t->data = data;
    pthread_create(&t->handle, 0, svo_shim_thread_entry, t);
    return t;
}
static inline void SDL_WaitThread(SDL_Thread *t, int *) { pthread_join(t->handle, 0); free(t); }
static inline SDL_mutex *SDL_CreateMutex() { SDL_mutex *m = (SDL_mutex *)malloc(sizeof(SDL_mutex)); pthread_mutex_init(&m->m, 0); return m; }
static inline void SDL_DestroyMutex(SDL_mutex *m) { pthread_mutex_destroy(&m->m); free(m); }
static inline int SDL_mutexP(SDL_mutex *m) { return pthread_mutex_lock(&m->m); }
static inline int SDL_mutexV(SDL_mutex *m) { return pthread_mutex_unlock(&m->m); }
static inline SDL_sem *SDL_CreateSemaphore(unsigned v) { SDL_sem *s = (SDL_sem *)malloc(sizeof(SDL_sem)); sem_init(&s->s, 0, v); return s; }
static inline void SDL_DestroySemaphore(SDL_sem *s) { sem_destroy(&s->s); free(s); }
static inline int SDL_SemWait(SDL_sem *s) { int r; while ((r = sem_wait(&s->s)) != 0) {} return r; }
static inline int SDL_SemPost(SDL_sem *s) { return sem_post(&s->s); }

/* ---- window ---- */
#define SDL_MUSTLOCK(s) 0

static inline int  SDL_Init(unsigned) { return 0; }
static inline void SDL_Quit() {}
static inline void SDL_WM_SetCaption(const char *, const char *) {}
static inline int  SDL_LockSurface(SDL_Surface *) { return 0; }
static inline void SDL_UnlockSurface(SDL_Surface *) {}
static inline void SDL_UpdateRect(SDL_Surface *s, int, int, unsigned, unsigned) { svo_shim_present(s); }
static inline SDL_Surface *SDL_SetVideoMode(int w, int h, int, unsigned) {
    SDL_Surface *s = (SDL_Surface *)malloc(sizeof(SDL_Surface));
    s->pixels = calloc((size_t)w*(size_t)h, 4);
    s->pitch = w*4;
    s->w = w;
    s->h = h;
    svo_shim_surface_created(s);
    return s;
}

#endif
