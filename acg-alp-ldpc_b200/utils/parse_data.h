// Matrix text I/O with the reference's names (utils/parse_data.h:6-54) plus the
// sparse ".rows" format this repo ships its matrices in.
#ifndef LDPC_B200_UTILS_PARSE_DATA_H
#define LDPC_B200_UTILS_PARSE_DATA_H

#include <sstream>

#include "codeword.h"

// Dense comma-separated 0/1 rows, one whitespace-separated token per row, optional
// trailing comma.  A cell is 1 iff the LAST character before its comma is '1'
// (so "2" reads as 0, as it does in the reference's G files), utils/parse_data.h:6-25.
inline TMatrix read_pcm(const string &filename) {
    ifstream in(filename);
    TMatrix rows;
    string token;
    bool cell = false;   // carried across rows like the reference's `t` (an empty cell repeats the last value)
    while (in >> token) {
        if (token.back() != ',') token.push_back(',');
        TCodeword row;
        for (char ch : token) {
            if (ch == ',') row.push_back(cell);
            else cell = (ch == '1');
        }
        rows.push_back(row);
    }
    return rows;
}

// "<count>\n<word>\n..." with INVERTED bits: '0' reads as true (utils/parse_data.h:28-42)
inline vector<TCodeword> read_codewords(const string &filename) {
    ifstream in(filename);
    int count = 0;
    in >> count;
    vector<TCodeword> words(max(count, 0));
    for (TCodeword &w : words) {
        string text;
        in >> text;
        w.resize(text.size());
        for (size_t i = 0; i < text.size(); ++i) w[i] = text[i] == '0';
    }
    return words;
}

// utils/parse_data.h:44-54: comma-separated, no trailing comma, one row per line
inline void save_matrix(const TMatrix &H, const string &filepath) {
    ofstream out(filepath);
    for (const TCodeword &row : H) {
        string line;
        line.reserve(2 * row.size());
        for (size_t i = 0; i < row.size(); ++i) {
            if (i) line.push_back(',');
            line.push_back(row[i] ? '1' : '0');
        }
        out << line << endl;
    }
}

// Sparse text: '#' comments, "m n", then per row "deg c0 c1 ..." (ascending columns).
inline TMatrix read_pcm_rows(const string &filename) {
    ifstream in(filename);
    string line;
    TMatrix H;
    size_t m = 0, n = 0;
    bool have_shape = false;
    while (getline(in, line)) {
        if (line.empty() || line[0] == '#') continue;
        istringstream ls(line);
        if (!have_shape) {
            ls >> m >> n;
            have_shape = true;
            H.reserve(m);
            continue;
        }
        size_t deg = 0;
        ls >> deg;
        TCodeword row(n, false);
        for (size_t k = 0; k < deg; ++k) {
            size_t col = 0;
            ls >> col;
            assert(col < n);
            row[col] = true;
        }
        H.push_back(row);
    }
    assert(H.size() == m);
    return H;
}

// data/<name>.txt (dense, what the reference's drivers open) when present, else data/<name>.rows
inline TMatrix load_matrix(const string &stem) {
    TMatrix H = read_pcm(stem + ".txt");
    if (!H.empty()) return H;
    return read_pcm_rows(stem + ".rows");
}

#endif
