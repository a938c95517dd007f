// host/sparse_matrix.hpp -- host CRS/COO containers (reference
// sparse_matrix.hpp:59-203), the MatrixMarket reader (sparse_matrix.hpp:225-350
// semantics: symmetric files are expanded entry-by-entry, then a STABLE sort by
// row only, so the within-row order is the file order), the
// COO->CRS conversion (utilities/utilities.hpp:326-367), the built-in synthetic
// generators and the device mirror handle.
#pragma once

#include "common.hpp"

#include <algorithm>
#include <cstring>
#include <fstream>
#include <numeric>
#include <sstream>

struct MatrixCRS {
    int n_rows{};
    int n_cols{};
    int nnz{};
    int *row_ptr = nullptr;
    int *col = nullptr;
    double *val = nullptr;

    MatrixCRS() = default;
    MatrixCRS(std::size_t num_rows, std::size_t num_cols, std::size_t num_nnz)
        : n_rows((int)num_rows), n_cols((int)num_cols), nnz((int)num_nnz) {
        row_ptr = new int[n_rows + 1];
        col = new int[nnz > 0 ? nnz : 1];
        val = new double[nnz > 0 ? nnz : 1];
    }
    MatrixCRS(const MatrixCRS &) = delete;
    MatrixCRS &operator=(const MatrixCRS &) = delete;
    ~MatrixCRS() {
        delete[] row_ptr;
        delete[] col;
        delete[] val;
    }
};

// Device mirror of a MatrixCRS (new in the build): an opaque C-ABI handle.
struct DeviceCRS {
    bis_context *dev = nullptr;
    bis_matrix *handle = nullptr;
    int64_t n_rows = 0, n_rows_global = 0, nnz = 0, nnz_global = 0;
    int n_levels = 0;
    DeviceCRS() = default;
    DeviceCRS(const DeviceCRS &) = delete;
    DeviceCRS &operator=(const DeviceCRS &) = delete;
    void refresh_info() {
        int64_t info[8];
        if (bis_matrix_info(handle, info) != 0) bis_fatal(bis_last_error());
        n_rows = info[0];
        n_rows_global = info[1];
        nnz = info[2];
        nnz_global = info[3];
        n_levels = (int)info[5];
    }
    ~DeviceCRS() {
        if (handle) bis_matrix_free(dev, handle);
    }
};

inline std::unique_ptr<DeviceCRS> upload_crs(bis_context *dev, const MatrixCRS *A) {
    auto d = std::make_unique<DeviceCRS>();
    d->dev = dev;
    if (bis_matrix_upload_crs(dev, A->n_rows, A->n_cols, A->nnz, A->row_ptr, A->col, A->val, &d->handle) != 0)
        bis_fatal(std::string("upload_crs: ") + bis_last_error());
    d->refresh_info();
    return d;
}
inline std::unique_ptr<DeviceCRS> upload_triangular(bis_context *dev, const MatrixCRS *T, bool upper) {
    auto d = std::make_unique<DeviceCRS>();
    d->dev = dev;
    if (bis_matrix_upload_triangular(dev, T->n_rows, T->nnz, T->row_ptr, T->col, T->val, upper ? 1 : 0,
                                     &d->handle) != 0)
        bis_fatal(std::string("upload_triangular: ") + bis_last_error());
    d->refresh_info();
    return d;
}

// takes ownership of a matrix the device library created (split / ILU(0) on the device)
inline std::unique_ptr<DeviceCRS> adopt_device_matrix(bis_context *dev, bis_matrix *handle) {
    auto d = std::make_unique<DeviceCRS>();
    d->dev = dev;
    d->handle = handle;
    d->refresh_info();
    return d;
}

struct MatrixCOO {
    long n_rows{};
    long n_cols{};
    long nnz{};
    bool is_sorted{};
    bool is_symmetric{};
    std::vector<int> I;
    std::vector<int> J;
    std::vector<double> values;

    // MatrixMarket "coordinate" reader: real / integer / pattern, general /
    // symmetric; throws std::runtime_error on I/O problems like the reference
    // (sparse_matrix.hpp:263-299).
    // sort_by_row = false keeps the file's order (the device sorts: bis_matrix_upload_coo)
    void read_from_mtx(const std::string &path, const bool sort_by_row = true) {
        std::ifstream f(path);
        if (!f) throw std::runtime_error("Unable to open file: " + path);
        std::string line;
        if (!std::getline(f, line)) throw std::runtime_error("Could not process Matrix Market banner in file: " + path);
        std::string banner, object, format, field, symmetry;
        {
            std::istringstream is(line);
            is >> banner >> object >> format >> field >> symmetry;
            for (auto *s : {&object, &format, &field, &symmetry})
                std::transform(s->begin(), s->end(), s->begin(), ::tolower);
        }
        if (banner != "%%MatrixMarket" || object != "matrix")
            throw std::runtime_error("Could not process Matrix Market banner in file: " + path);
        const bool pattern = field == "pattern";
        const bool symm = symmetry == "symmetric";
        if (format != "coordinate" || !(field == "real" || field == "integer" || pattern) ||
            !(symm || symmetry == "general"))
            throw std::runtime_error("Unsupported matrix format in file: " + path);
        while (std::getline(f, line))
            if (!line.empty() && line[0] != '%') break;
        long nr = 0, nc = 0, nz = 0;
        {
            std::istringstream is(line);
            if (!(is >> nr >> nc >> nz)) throw std::runtime_error("Error reading matrix from file: " + path);
        }
        if (nr != nc) throw std::runtime_error("Matrix must be square.");
        std::vector<int> rows, cols;
        std::vector<double> vals;
        rows.reserve(symm ? 2 * nz : nz);
        cols.reserve(symm ? 2 * nz : nz);
        vals.reserve(symm ? 2 * nz : nz);
        for (long k = 0; k < nz; ++k) {
            long i = 0, j = 0;
            double v = 1.0;
            if (!(f >> i >> j)) throw std::runtime_error("Error reading matrix from file: " + path);
            if (!pattern && !(f >> v)) throw std::runtime_error("Error reading matrix from file: " + path);
            rows.push_back((int)(i - 1));
            cols.push_back((int)(j - 1));
            vals.push_back(v);
            if (symm && i != j) {
                rows.push_back((int)(j - 1));
                cols.push_back((int)(i - 1));
                vals.push_back(v);
            }
        }
        const size_t m = vals.size();
        std::vector<size_t> perm(m);
        std::iota(perm.begin(), perm.end(), 0);
        if (sort_by_row) std::stable_sort(perm.begin(), perm.end(), [&](size_t a, size_t b) { return rows[a] < rows[b]; });
        I.resize(m);
        J.resize(m);
        values.resize(m);
        for (size_t k = 0; k < m; ++k) {
            I[k] = rows[perm[k]];
            J[k] = cols[perm[k]];
            values[k] = vals[perm[k]];
        }
        n_rows = nr;
        n_cols = nc;
        nnz = (long)m;
        is_sorted = sort_by_row;
        is_symmetric = false;
    }
};

// utilities/utilities.hpp:326-367: entries are already grouped by row
inline void convert_coo_to_crs(const MatrixCOO *coo, MatrixCRS *crs) {
    delete[] crs->row_ptr;
    delete[] crs->col;
    delete[] crs->val;
    crs->n_rows = (int)coo->n_rows;
    crs->n_cols = (int)coo->n_cols;
    crs->nnz = (int)coo->nnz;
    crs->row_ptr = new int[crs->n_rows + 1]();
    crs->col = new int[crs->nnz > 0 ? crs->nnz : 1];
    crs->val = new double[crs->nnz > 0 ? crs->nnz : 1];
    for (long k = 0; k < coo->nnz; ++k) {
        crs->col[k] = coo->J[k];
        crs->val[k] = coo->values[k];
        ++crs->row_ptr[coo->I[k] + 1];
    }
    for (int r = 0; r < crs->n_rows; ++r) crs->row_ptr[r + 1] += crs->row_ptr[r];
    if (crs->row_ptr[crs->n_rows] != crs->nnz) bis_fatal("ERROR: converting to CRS.");
}

// ---- synthetic matrices (host versions; the device generators in
// csrc/bis_matrix.cu produce the same arrays) --------------------------------------
inline std::unique_ptr<MatrixCRS> generate_hpcg(int nx, int ny, int nz) {
    const long long n = 1LL * nx * ny * nz;
    const long long nnz = 1LL * (nx > 1 ? 3 * nx - 2 : 1) * (ny > 1 ? 3 * ny - 2 : 1) * (nz > 1 ? 3 * nz - 2 : 1);
    if (n >= INT32_MAX || nnz >= INT32_MAX)
        bis_fatal("generate_hpcg: matrix exceeds the host MatrixCRS's 32-bit nnz (use the device generator)");
    auto A = std::make_unique<MatrixCRS>((size_t)n, (size_t)n, (size_t)nnz);
    int k = 0;
    A->row_ptr[0] = 0;
    for (int z = 0; z < nz; ++z)
        for (int y = 0; y < ny; ++y)
            for (int x = 0; x < nx; ++x) {
                const int row = (z * ny + y) * nx + x;
                for (int dz = -1; dz <= 1; ++dz) {
                    if (z + dz < 0 || z + dz >= nz) continue;
                    for (int dy = -1; dy <= 1; ++dy) {
                        if (y + dy < 0 || y + dy >= ny) continue;
                        for (int dx = -1; dx <= 1; ++dx) {
                            if (x + dx < 0 || x + dx >= nx) continue;
                            A->col[k] = row + (dz * ny + dy) * nx + dx;
                            A->val[k] = (dx == 0 && dy == 0 && dz == 0) ? 26.0 : -1.0;
                            ++k;
                        }
                    }
                }
                A->row_ptr[row + 1] = k;
            }
    return A;
}

inline double splitmix_unit(uint64_t seed, uint64_t idx) {
    uint64_t z = seed + (idx + 1) * 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z ^= z >> 31;
    return (double)(z >> 11) * (1.0 / 9007199254740992.0);
}

inline std::unique_ptr<MatrixCRS> generate_anderson(int lx, int ly, int lz, double ranpot, double t,
                                                    uint64_t seed, bool periodic) {
    const long long n = 1LL * lx * ly * lz;
    if (7 * n >= INT32_MAX) bis_fatal("generate_anderson: lattice too large for the host MatrixCRS");
    std::vector<int> rp(n + 1, 0), col;
    std::vector<double> val;
    col.reserve(7 * n);
    val.reserve(7 * n);
    const int dxs[7] = {0, 0, -1, 0, 1, 0, 0}, dys[7] = {0, -1, 0, 0, 0, 1, 0}, dzs[7] = {-1, 0, 0, 0, 0, 0, 1};
    for (long long row = 0; row < n; ++row) {
        const long long x = row % lx, y = (row / lx) % ly, z = row / (1LL * lx * ly);
        std::vector<std::pair<long long, double>> ent;
        for (int k = 0; k < 7; ++k) {
            long long xx = x + dxs[k], yy = y + dys[k], zz = z + dzs[k];
            bool ok = true;
            if (periodic) {
                if (dxs[k] && lx <= 2) ok = ok && xx >= 0 && xx < lx;
                if (dys[k] && ly <= 2) ok = ok && yy >= 0 && yy < ly;
                if (dzs[k] && lz <= 2) ok = ok && zz >= 0 && zz < lz;
                xx = (xx + lx) % lx;
                yy = (yy + ly) % ly;
                zz = (zz + lz) % lz;
            } else {
                ok = xx >= 0 && xx < lx && yy >= 0 && yy < ly && zz >= 0 && zz < lz;
            }
            if (!ok) continue;
            double u2 = 2.0 * splitmix_unit(seed, (uint64_t)row);
            double v = (k == 3) ? ranpot * (u2 - 1.0) : -t;
            ent.push_back({(zz * ly + yy) * lx + xx, v});
        }
        std::stable_sort(ent.begin(), ent.end(), [](auto &a, auto &b) { return a.first < b.first; });
        for (auto &e : ent) {
            col.push_back((int)e.first);
            val.push_back(e.second);
        }
        rp[row + 1] = (int)col.size();
    }
    auto A = std::make_unique<MatrixCRS>((size_t)n, (size_t)n, col.size());
    std::memcpy(A->row_ptr, rp.data(), sizeof(int) * (n + 1));
    std::memcpy(A->col, col.data(), sizeof(int) * col.size());
    std::memcpy(A->val, val.data(), sizeof(double) * val.size());
    return A;
}

// Matrix "names" understood at run time (the reference picks file vs SCAMAC at
// compile time, main.cpp:48-54): HPCG-<n>, HPCG-<nx>-<ny>-<nz>,
// Anderson,Lx=..,Ly=..,Lz=..,ranpot=..[,t=..,seed=..,boundary_conditions=open|periodic];
// anything else is a MatrixMarket path.
struct MatrixSpec {
    enum Kind { File, Hpcg, Anderson } kind = File;
    int nx = 0, ny = 0, nz = 0;
    double ranpot = 1.0, t = 1.0;
    uint64_t seed = 1;
    bool periodic = false;
    std::string path;
};

inline MatrixSpec parse_matrix_spec(const std::string &name) {
    MatrixSpec s;
    s.path = name;
    auto lower = name;
    std::transform(lower.begin(), lower.end(), lower.begin(), ::tolower);
    if (lower.rfind("hpcg-", 0) == 0 && lower.find(".mtx") == std::string::npos) {
        int a = 0, b = 0, c = 0;
        int got = std::sscanf(lower.c_str(), "hpcg-%d-%d-%d", &a, &b, &c);
        if (got == 1) b = c = a;
        if ((got == 1 || got == 3) && a > 0 && b > 0 && c > 0) {
            s.kind = MatrixSpec::Hpcg;
            s.nx = a; s.ny = b; s.nz = c;
        }
        return s;
    }
    if (lower.rfind("anderson", 0) == 0 && lower.find(".mtx") == std::string::npos) {
        s.kind = MatrixSpec::Anderson;
        s.nx = s.ny = s.nz = 5;
        std::istringstream is(lower);
        std::string tok;
        while (std::getline(is, tok, ',')) {
            auto eq = tok.find('=');
            if (eq == std::string::npos) continue;
            std::string k = tok.substr(0, eq), v = tok.substr(eq + 1);
            if (k == "lx") s.nx = std::atoi(v.c_str());
            else if (k == "ly") s.ny = std::atoi(v.c_str());
            else if (k == "lz") s.nz = std::atoi(v.c_str());
            else if (k == "ranpot") s.ranpot = std::atof(v.c_str());
            else if (k == "t") s.t = std::atof(v.c_str());
            else if (k == "seed") s.seed = std::strtoull(v.c_str(), nullptr, 10);
            else if (k == "boundary_conditions") s.periodic = (v == "periodic");
        }
        return s;
    }
    return s;
}
