#include "hierarchy.h"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <limits>
#include <cstdlib>
#include <queue>
#include <thread>

namespace gmg {

// Worker threads of the per-point selection: GMG_HIERARCHY_THREADS, else the hardware concurrency (at most 8).
static int hierarchy_threads() {
    if (const char* e = std::getenv("GMG_HIERARCHY_THREADS")) return std::max(1, std::atoi(e));
    return (int)std::max(1u, std::thread::hardware_concurrency());
}

namespace {

struct Vec3 {
    double x, y, z;
};
inline Vec3 operator+(const Vec3& a, const Vec3& b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline Vec3 operator-(const Vec3& a, const Vec3& b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline Vec3 operator*(double s, const Vec3& a) { return {s * a.x, s * a.y, s * a.z}; }
inline double dot(const Vec3& a, const Vec3& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline Vec3 cross(const Vec3& a, const Vec3& b) {
    return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
inline double norm(const Vec3& a) { return std::sqrt(dot(a, a)); }
inline Vec3 normalized(const Vec3& a) {
    const double n2 = dot(a, a);
    return n2 > 0.0 ? (1.0 / std::sqrt(n2)) * a : a;
}

struct Points {
    std::vector<double> xyz;
    int64_t n = 0;
    Vec3 at(int64_t i) const { return {xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]}; }
    void set(int64_t i, const Vec3& v) {
        xyz[3 * i] = v.x;
        xyz[3 * i + 1] = v.y;
        xyz[3 * i + 2] = v.z;
    }
};

struct NeighArray {  // n x width, -1 padded
    std::vector<int> a;
    int64_t n = 0;
    int width = 0;
    int at(int64_t i, int j) const { return a[i * width + j]; }
};

// fn(lo, hi) on contiguous chunks of [0, n), one chunk per worker thread (results must not depend on the split).
template <typename F>
void parallel_chunks(int64_t n, int64_t min_chunk, F&& fn) {
    const int n_threads = (int)std::max<int64_t>(1, std::min<int64_t>({(int64_t)hierarchy_threads(), 16, n / std::max<int64_t>(min_chunk, 1) + 1}));
    if (n_threads == 1) {
        fn((int64_t)0, n);
        return;
    }
    std::vector<std::thread> pool;
    for (int t = 1; t < n_threads; ++t) pool.emplace_back([&, t] { fn(n * t / n_threads, n * (t + 1) / n_threads); });
    fn((int64_t)0, n / n_threads);
    for (auto& th : pool) th.join();
}

struct Clock {
    std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
    double ms() const { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count(); }
};

// Mean length of the non-degenerate stored edges (ms.cpp:695-711).
double average_edge_length(const Points& pos, const NeighArray& ng) {
    double sum = 0.0;
    int64_t count = 0;
    for (int64_t i = 0; i < pos.n; ++i) {
        const Vec3 p = pos.at(i);
        for (int j = 0; j < ng.width; ++j) {
            const int q = ng.at(i, j);
            if (q < 0) continue;
            const double d = norm(p - pos.at(q));
            if (d > 0) {
                sum += d;
                ++count;
            }
        }
    }
    return sum / (double)count;
}

// Greedy index-order disk sampling over one- and two-rings (ms.cpp:975-1013).
std::vector<int> fast_disk_sample(const Points& pos, const NeighArray& ng, double radius,
                                  std::vector<double>& dist, std::vector<int>& nearest) {
    std::vector<char> visited(pos.n, 0);
    std::vector<int> picked;
    int sample = 0;
    for (int64_t i = 0; i < pos.n; ++i) {
        if (visited[i]) continue;
        picked.push_back((int)i);
        nearest[i] = sample;
        const Vec3 pi = pos.at(i);
        for (int j = 0; j < ng.width; ++j) {
            const int a = ng.at(i, j);
            if (a < 0) break;
            const Vec3 pa = pos.at(a);
            const double d1 = norm(pi - pa);
            if (!(d1 < radius)) continue;
            visited[a] = 1;
            if (d1 < dist[a]) {
                dist[a] = d1;
                nearest[a] = sample;
            }
            for (int j2 = 0; j2 < ng.width; ++j2) {
                const int b = ng.at(a, j2);
                if (b < 0) break;
                const double d2 = d1 + norm(pa - pos.at(b));
                if (d2 < radius) {
                    visited[b] = 1;
                    if (d2 < dist[b]) {
                        dist[b] = d2;
                        nearest[b] = sample;
                    }
                }
            }
        }
        ++sample;
    }
    return picked;
}

struct QueueItem {  // ordering on distance only, like the reference's VertexPair (sampling.h:8-13)
    int v;
    double d;
    bool operator>(const QueueItem& o) const { return d > o.d; }
};

// Multi-source Dijkstra seeded with the distances left by the sampler (ms.cpp:1015-1056).
void cluster_by_dijkstra(const Points& pos, const std::vector<int>& sources, const NeighArray& ng,
                         std::vector<double>& dist, std::vector<int>& nearest) {
    std::priority_queue<QueueItem, std::vector<QueueItem>, std::greater<QueueItem>> queue;
    for (size_t s = 0; s < sources.size(); ++s) {
        dist[sources[s]] = 0.0;
        queue.push({sources[s], 0.0});
        nearest[sources[s]] = (int)s;
    }
    while (!queue.empty()) {
        const QueueItem top = queue.top();
        const int owner = nearest[top.v];
        const Vec3 p = pos.at(top.v);
        queue.pop();
        for (int j = 0; j < ng.width; ++j) {
            const int q = ng.at(top.v, j);
            if (q < 0) continue;
            const double cand = top.d + norm(pos.at(q) - p);
            if (cand < dist[q]) {
                dist[q] = cand;
                queue.push({q, cand});
                nearest[q] = owner;
            }
        }
    }
}

struct EdgeFlag {  // the reference keeps these in a std::map<int, float>
    int key;
    float value;
};

inline EdgeFlag* find_flag(std::vector<EdgeFlag>& flags, int key) {
    for (auto& f : flags)
        if (f.key == key) return &f;
    return nullptr;
}

inline void set_flag(std::vector<EdgeFlag>& flags, int key, float value) {
    if (EdgeFlag* f = find_flag(flags, key))
        f->value = value;
    else
        flags.push_back({key, value});
}

// Projected barycentric test with the edge side effects of ms.cpp:471-507.
// Returns |distance to the triangle plane| when the projection lies inside, else -1.
double in_triangle(const Vec3& p, const int tri[3], const Vec3& nrm, const Points& pos, double bary[3],
                   std::vector<EdgeFlag>& flags) {
    const Vec3 v1 = pos.at(tri[0]), v2 = pos.at(tri[1]), v3 = pos.at(tri[2]);
    const Vec3 v1p = p - v1;
    const Vec3 e12 = v2 - v1;
    const Vec3 e13 = v3 - v1;
    const double d = dot(v1p, nrm);
    const Vec3 proj = p - d * nrm;
    const double dbl_area = dot(cross(e12, e13), nrm);
    bary[0] = dot(cross(v3 - v2, proj - v2), nrm) / dbl_area;
    bary[1] = dot(cross(v1 - v3, proj - v3), nrm) / dbl_area;
    bary[2] = 1.0 - bary[0] - bary[1];
    if (!find_flag(flags, tri[1])) flags.push_back({tri[1], (float)norm(v1p - dot(v1p, e12) * e12)});
    if (!find_flag(flags, tri[2])) flags.push_back({tri[2], (float)norm(v1p - dot(v1p, e13) * e13)});
    if (bary[0] < 0.0 || bary[1] < 0.0) set_flag(flags, tri[1], -1.0f);
    if (bary[0] < 0.0 || bary[2] < 0.0) set_flag(flags, tri[2], -1.0f);
    if (bary[0] >= 0.0 && bary[1] >= 0.0 && bary[2] >= 0.0) return std::fabs(d);
    return -1.0;
}

void inverse_distance_weights(const Points& pos, const Vec3& p, const int* ids, int count, double* w) {
    double total = 0.0;
    for (int j = 0; j < count; ++j) {
        w[j] = 1.0 / std::max(1e-8, norm(p - pos.at(ids[j])));
        total += w[j];
    }
    for (int j = 0; j < count; ++j) w[j] /= total;
}

// Clamped parameter of p along the segment c -> q (ms.cpp:320-326, 397-402).
inline double edge_parameter(const Vec3& p, const Vec3& c, const Vec3& q) {
    const Vec3 e = q - c;
    const double len = std::max(norm(e), 1e-8);
    const double w = dot(p - c, normalized(e)) / len;
    return std::min(std::max(w, 0.0), 1.0);
}

struct RowBuilder {  // up to three weighted entries per fine point
    std::vector<int> indptr{0};
    std::vector<int> cols;
    std::vector<double> vals;
    void push_row(const int* c, const double* w, int count) {
        int order[3] = {0, 1, 2};
        for (int a = 1; a < count; ++a)  // insertion sort of at most three entries by column
            for (int b = a; b > 0 && c[order[b]] < c[order[b - 1]]; --b) std::swap(order[b], order[b - 1]);
        for (int t = 0; t < count;) {
            int u = t;
            double s = 0.0;
            while (u < count && c[order[u]] == c[order[t]]) s += w[order[u++]];
            cols.push_back(c[order[t]]);
            vals.push_back(s);
            t = u;
        }
        indptr.push_back((int)cols.size());
    }
};

}  // namespace

void build_hierarchy(const double* pos_in, int64_t n, const int* neigh_in, int kn, const HierarchyOptions& opt,
                     Hierarchy& out) {
    Clock total;
    out = Hierarchy();
    auto& tm = out.timing;
    tm["n_vertices"] = (double)n;
    for (const char* key : {"PDS", "sampling", "cluster", "next_neighborhood", "next_positions", "triangle_finding",
                            "triangle_selection"})
        tm[key] = 0.0;

    Points level;
    level.n = n;
    level.xyz.assign(pos_in, pos_in + 3 * n);
    NeighArray ng;
    ng.n = n;
    ng.width = kn;
    ng.a.assign(neigh_in, neigh_in + n * (int64_t)kn);

    out.dof.push_back(n);
    int k = 0;
    while (level.n > opt.low_bound && k < 10) {
        const int64_t nf = level.n;
        const double radius = std::cbrt(opt.ratio) * average_edge_length(level, ng);

        // -- sampling
        Clock t_sample;
        std::vector<double> dist(nf, std::numeric_limits<double>::max());
        std::vector<int> nearest(nf, 0);
        std::vector<int> picked = fast_disk_sample(level, ng, radius, dist, nearest);
        if ((int64_t)picked.size() < opt.low_bound) break;
        const int64_t nc = (int64_t)picked.size();
        tm["sampling"] += t_sample.ms();
        if (opt.verbose) std::printf("level %d: %lld -> %lld points\n", k, (long long)nf, (long long)nc);

        // -- graph-Voronoi clustering
        Clock t_cluster;
        cluster_by_dijkstra(level, picked, ng, dist, nearest);
        tm["cluster"] += t_cluster.ms();

        // -- coarse adjacency: clusters joined by a fine edge (ms.cpp:178-207)
        Clock t_neigh;
        // (every cluster's neighbour set is sorted and made unique, so it does not depend on the order its fine
        // edges are visited in: clusters are processed on worker threads from a members-by-cluster list)
        std::vector<std::vector<int>> adj(nc);
        std::vector<int64_t> member_ptr((size_t)nc + 1, 0);
        for (int64_t i = 0; i < nf; ++i) ++member_ptr[(size_t)nearest[i] + 1];
        for (int64_t c = 0; c < nc; ++c) member_ptr[c + 1] += member_ptr[c];
        std::vector<int> members_of((size_t)nf);
        {
            std::vector<int64_t> at(member_ptr.begin(), member_ptr.end() - 1);
            for (int64_t i = 0; i < nf; ++i) members_of[at[nearest[i]]++] = (int)i;
        }
        parallel_chunks(nc, 2048, [&](int64_t c_lo, int64_t c_hi) {
            for (int64_t c = c_lo; c < c_hi; ++c) {
                std::vector<int>& a = adj[c];
                for (int64_t t = member_ptr[c]; t < member_ptr[c + 1]; ++t) {
                    const int i = members_of[t];
                    for (int j = 0; j < ng.width; ++j) {
                        const int q = ng.at(i, j);
                        if (q < 0) break;
                        if (nearest[q] != c) a.push_back(nearest[q]);
                    }
                }
                std::sort(a.begin(), a.end());
                a.erase(std::unique(a.begin(), a.end()), a.end());
            }
        });
        size_t widest = 0;
        for (const auto& a : adj) widest = std::max(widest, a.size());
        NeighArray ng_next;
        ng_next.n = nc;
        ng_next.width = (int)std::max<size_t>(widest, 1);
        ng_next.a.assign(nc * (int64_t)ng_next.width, -1);
        for (int64_t c = 0; c < nc; ++c) {
            ng_next.a[c * ng_next.width] = (int)c;
            int slot = 1;
            for (int q : adj[c]) {
                if (q == c) continue;
                if (slot >= (int)widest) break;  // the widest row loses its last neighbour, as upstream
                ng_next.a[c * ng_next.width + slot++] = q;
            }
        }
        tm["next_neighborhood"] += t_neigh.ms();

        // -- coarse positions (ms.cpp:216-240)
        Clock t_pos;
        Points coarse;
        coarse.n = nc;
        coarse.xyz.assign(3 * nc, 0.0);
        if (opt.nested) {
            for (int64_t c = 0; c < nc; ++c) coarse.set(c, level.at(picked[c]));
        } else {
            std::vector<int> members(nc, 0);
            for (int64_t i = 0; i < nf; ++i) {
                const int c = nearest[i];
                coarse.set(c, coarse.at(c) + level.at(i));
                ++members[c];
            }
            for (int64_t c = 0; c < nc; ++c) {
                if (members[c] == 1) {
                    Vec3 s = level.at(picked[c]);
                    for (int q : adj[c]) s = s + level.at(picked[q]);
                    coarse.set(c, (1.0 / (adj[c].size() + 1.0)) * s);
                } else {
                    const Vec3 s = coarse.at(c);
                    const double m = (double)members[c];
                    coarse.set(c, {s.x / m, s.y / m, s.z / m});
                }
            }
        }
        if (opt.debug) out.level_points.push_back(coarse.xyz);
        tm["next_positions"] += t_pos.ms();

        // -- candidate triangles from the Voronoi dual (ms.cpp:248-281)
        Clock t_tri;
        // (the triangles a coarse point creates depend on nothing but the adjacency: they are listed per point on
        // worker threads and numbered afterwards in point order, the order the sequential loop creates them in)
        std::vector<int> tris;
        std::vector<Vec3> tri_normals;
        std::vector<std::vector<int>> incident(nc);
        {
            std::vector<std::vector<int>> pairs_of(nc);  // (v2, v3) of the triangles created by point c
            parallel_chunks(nc, 2048, [&](int64_t c_lo, int64_t c_hi) {
                for (int64_t c = c_lo; c < c_hi; ++c) {
                    const auto& a = adj[c];
                    for (size_t s2 = 0; s2 < a.size(); ++s2) {
                        const int v2 = a[s2];
                        if (v2 < c) continue;
                        for (size_t t = s2 + 1; t < a.size(); ++t) {
                            const int v3 = a[t];
                            if (v3 < c) continue;
                            if (opt.check_voronoi && !std::binary_search(adj[v2].begin(), adj[v2].end(), v3)) continue;
                            pairs_of[c].push_back(v2);
                            pairs_of[c].push_back(v3);
                        }
                    }
                }
            });
            std::vector<int64_t> first((size_t)nc + 1, 0);
            for (int64_t c = 0; c < nc; ++c) first[c + 1] = first[c] + (int64_t)pairs_of[c].size() / 2;
            const int64_t n_tri = first[nc];
            tris.resize(3 * (size_t)n_tri);
            tri_normals.resize((size_t)n_tri);
            parallel_chunks(nc, 2048, [&](int64_t c_lo, int64_t c_hi) {
                for (int64_t c = c_lo; c < c_hi; ++c)
                    for (size_t t = 0; t < pairs_of[c].size() / 2; ++t) {
                        const int64_t id = first[c] + (int64_t)t;
                        const int v2 = pairs_of[c][2 * t], v3 = pairs_of[c][2 * t + 1];
                        tris[3 * id] = (int)c, tris[3 * id + 1] = v2, tris[3 * id + 2] = v3;
                        tri_normals[id] = normalized(cross(coarse.at(v2) - coarse.at(c), coarse.at(v3) - coarse.at(c)));
                    }
            });
            for (int64_t id = 0; id < n_tri; ++id)  // ascending triangle id per point, as the sequential loop leaves them
                for (int v = 0; v < 3; ++v) incident[tris[3 * id + v]].push_back((int)id);
        }
        if (opt.debug) out.all_triangles.push_back(tris);
        tm["triangle_finding"] += t_tri.ms();

        // -- prolongation weights of every fine point (ms.cpp:293-453)
        Clock t_sel;
        RowBuilder rows;
        rows.indptr.reserve(nf + 1);
        rows.cols.reserve(3 * nf);
        rows.vals.reserve(3 * nf);
        std::vector<int> missing;
        if (opt.debug) missing.assign(nf, 0);
        // Every fine point is independent of the others (multigrid_solver.cpp:293-453): the selection runs on
        // worker threads over contiguous chunks into per-point slots, the rows are assembled in order afterwards.
        std::vector<int> sel_ids(3 * (size_t)nf);
        std::vector<double> sel_w(3 * (size_t)nf);
        std::vector<unsigned char> sel_count((size_t)nf);
        auto select = [&](int64_t i, int* ids, double* w, std::vector<EdgeFlag>& flags,
                          std::vector<std::pair<double, int>>& by_distance) -> int {
            const Vec3 p = level.at(i);
            const int c = nearest[i];
            const Vec3 pc = coarse.at(c);
            if (opt.nested && picked[c] == i) {
                ids[0] = c, w[0] = 1.0;
                return 1;
            }
            const auto& a = adj[c];
            if (a.empty()) {
                ids[0] = c, w[0] = 1.0;
                return 1;
            }
            if (a.size() == 1) {
                ids[0] = c, ids[1] = a[0];
                if (opt.weighting == BARYCENTRIC) {
                    w[1] = edge_parameter(p, pc, coarse.at(a[0]));
                    w[0] = 1.0 - w[1];
                } else if (opt.weighting == UNIFORM) {
                    w[0] = w[1] = 0.5;
                } else {
                    inverse_distance_weights(coarse, p, ids, 2, w);
                }
                return 2;
            }
            // first incident triangle (creation order, rotated so c leads) containing the projection
            flags.clear();
            bool found = false;
            int hit[3] = {0, 0, 0};
            double hit_bary[3] = {0, 0, 0};
            for (int id : incident[c]) {
                int tri[3] = {tris[3 * id], tris[3 * id + 1], tris[3 * id + 2]};
                while (tri[0] != c) std::rotate(tri, tri + 1, tri + 3);
                double bary[3];
                const double d = in_triangle(p, tri, tri_normals[id], coarse, bary, flags);
                if (d >= 0.0 && d < std::numeric_limits<double>::max()) {
                    found = true;
                    std::copy(tri, tri + 3, hit);
                    std::copy(bary, bary + 3, hit_bary);
                    break;
                }
            }
            if (found) {
                if (opt.weighting == BARYCENTRIC)
                    std::copy(hit_bary, hit_bary + 3, w);
                else if (opt.weighting == UNIFORM)
                    w[0] = w[1] = w[2] = 1.0 / 3.0;
                else
                    inverse_distance_weights(coarse, p, hit, 3, w);
                std::copy(hit, hit + 3, ids);
                return 3;
            }
            if (opt.debug) missing[i] = 1;
            // else the lowest-numbered neighbour whose edge is still flagged "inside"
            int edge_to = -1;
            for (const auto& f : flags)
                if (f.value >= 0.0f && (edge_to < 0 || f.key < edge_to)) edge_to = f.key;
            if (edge_to >= 0) {
                ids[0] = c, ids[1] = edge_to;
                if (opt.weighting == BARYCENTRIC) {
                    w[1] = edge_parameter(p, pc, coarse.at(edge_to));
                    w[0] = 1.0 - w[1];
                } else if (opt.weighting == UNIFORM) {
                    w[0] = w[1] = 0.5;
                } else {
                    inverse_distance_weights(coarse, p, ids, 2, w);
                }
                return 2;
            }
            // else own cluster plus the two closest stored coarse neighbours, inverse-distance weighted
            by_distance.clear();
            for (int j = 0; j < ng_next.width; ++j) {
                const int q = ng_next.at(c, j);
                if (q < 0 || q == c) continue;
                by_distance.emplace_back(norm(p - coarse.at(q)), q);
            }
            std::sort(by_distance.begin(), by_distance.end(),
                      [](const std::pair<double, int>& x, const std::pair<double, int>& y) { return x.first < y.first; });
            int count = 1;
            ids[0] = c;
            for (size_t j = 0; j < by_distance.size() && count < 3; ++j) ids[count++] = by_distance[j].second;
            inverse_distance_weights(coarse, p, ids, count, w);
            return count;
        };
        {
            const int n_threads = (int)std::max<int64_t>(1, std::min<int64_t>({(int64_t)hierarchy_threads(), 8, nf / 4096 + 1}));
            auto work = [&](int t) {
                std::vector<EdgeFlag> flags;
                std::vector<std::pair<double, int>> by_distance;
                const int64_t lo = nf * t / n_threads, hi = nf * (t + 1) / n_threads;
                for (int64_t i = lo; i < hi; ++i)
                    sel_count[i] = (unsigned char)select(i, &sel_ids[3 * i], &sel_w[3 * i], flags, by_distance);
            };
            std::vector<std::thread> pool;
            for (int t = 1; t < n_threads; ++t) pool.emplace_back(work, t);
            work(0);
            for (auto& th : pool) th.join();
        }
        for (int64_t i = 0; i < nf; ++i) rows.push_row(&sel_ids[3 * i], &sel_w[3 * i], sel_count[i]);
        if (opt.debug) out.no_tri_found.push_back(std::move(missing));
        tm["triangle_selection"] += t_sel.ms();

        HostCsr u;
        u.rows = nf;
        u.cols = nc;
        u.indptr.swap(rows.indptr);
        u.indices.swap(rows.cols);
        u.data.swap(rows.vals);
        out.U.push_back(std::move(u));
        out.samples.push_back(std::move(picked));
        out.nearest_source.push_back(std::move(nearest));
        out.dof.push_back(nc);

        level = std::move(coarse);
        ng = std::move(ng_next);
        ++k;
    }
    tm["levels"] = (double)out.U.size();
    tm["hierarchy"] = total.ms();
}

}  // namespace gmg
