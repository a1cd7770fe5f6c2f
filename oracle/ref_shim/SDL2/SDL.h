/* Minimal stand-in for <SDL2/SDL.h>: only the typedefs and the two structs the reference's
 * Material / MaterialInstance / Tiles / Particle / world.cpp tick lines touch.
 * Test infrastructure for oracle/_ref (see oracle/README.md); not part of the product. */
#ifndef FSS_REF_SDL_H
#define FSS_REF_SDL_H
#include <stdint.h>
typedef uint8_t Uint8;
typedef uint16_t Uint16;
typedef uint32_t Uint32;
typedef int32_t Sint32;
typedef struct SDL_Rect {
    int x, y, w, h;
} SDL_Rect;
typedef struct SDL_Surface {
    int w, h, pitch;
    void* pixels;
} SDL_Surface;
#endif
