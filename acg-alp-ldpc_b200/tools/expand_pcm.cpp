// data/<name>.rows (sparse index lists) -> data/<name>.txt (the dense comma-separated
// layout read_pcm() and the reference's drivers open).   usage: expand_pcm in.rows out.txt
#include "../utils/parse_data.h"

int main(int argc, char **argv) {
    if (argc != 3) {
        cerr << "usage: expand_pcm <in.rows> <out.txt>" << endl;
        return 2;
    }
    TMatrix H = read_pcm_rows(argv[1]);
    if (H.empty()) return 1;
    save_matrix(H, argv[2]);
    return 0;
}
