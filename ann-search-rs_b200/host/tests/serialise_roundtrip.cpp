// Cross-language check of the saved-index mirror: loads an index.bin written by python/annb200/serialise.py with
// host/annb200_serialise.hpp and writes it back; the Python test compares the two files byte for byte.
//   serialise_roundtrip exhaustive|ivf <dir_in> <dir_out>     -> prints "OK"
//   serialise_roundtrip variant exhaustive|ivf <dir_in>       -> prints the error variant of the load (or "OK")
#include <cstdio>
#include <string>

#include "../annb200_serialise.hpp"

int main(int argc, char** argv) {
    using namespace annb200::serialise;
    if (argc < 4) { std::printf("usage\n"); return 2; }
    const std::string mode = argv[1];
    try {
        if (mode == "variant") {
            const std::string kind = argv[2];
            if (kind == "exhaustive") (void)load_exhaustive(argv[3]);
            else (void)load_ivf(argv[3]);
            std::printf("OK\n");
            return 0;
        }
        if (mode == "exhaustive") save_exhaustive(argv[3], load_exhaustive(argv[2]));
        else save_ivf(argv[3], load_ivf(argv[2]));
        std::printf("OK\n");
        return 0;
    } catch (const SerialiseError& e) {
        std::printf("%s\n", e.variant().c_str());
        return mode == "variant" ? 0 : 1;
    }
}
