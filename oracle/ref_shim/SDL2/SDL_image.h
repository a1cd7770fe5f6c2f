/* empty stand-in: Textures.hpp includes it, nothing on the tick path uses it */
