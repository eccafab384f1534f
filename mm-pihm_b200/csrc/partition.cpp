// partition.cpp -- mesh partitioning for the multi-GPU RHS (host code, no CUDA).
//
// The reference has no distributed mode; the coupling of its RHS is
// strictly local (element <-> 3 edge neighbours, element <-> adjacent river
// segment, river <-> downstream segment), so the mesh shards with
// one halo exchange per RHS evaluation.
//
// Partition p owns a contiguous chunk of the locality ordering (reorder.h), listed as
// [interior | boundary] (see below), and every river segment whose left-bank element it owns.  Its LOCAL mesh is again
// a plain pihm_b200_mesh (same column tables, local 1-based numbering):
//     elements = [ owned | ghosts grouped by owner rank ]
//     rivers   = [ owned | ghosts grouped by owner rank ]
// Ghost elements are everything an exact RHS of the owned entities reads:
//   * two rings of edge neighbours (|grad h| of a neighbour needs the heads of
//     ITS neighbours, lat_flow.c:33-36,134-170),
//   * both banks of every river that is computed locally (owned rivers, rivers
//     touching owned elements, upstream segments of owned rivers -- their fluxes
//     feed RiverToElem and the upstream sums, river_flow.c:94-181) and of the
//     downstream segments of those (SubFlowRiverToRiver, river_flow.c:47-57).
// Ghost rivers additionally include segments that only lend their stage to the
// friction slope of a ghost element (lat_flow.c:141-151).  Links that leave the
// local set are closed (boundary edge / outlet code): they are only ever reached
// from ghost entities whose results are discarded.
#include <algorithm>
#include <cstring>
#include <map>
#include <numeric>
#include <set>
#include <string>
#include <vector>
#include "pihm_b200.h"
#include "reorder.h"

namespace pb { void set_error(const std::string &msg); }

struct pihm_b200_partition {
    int nparts = 0, ne = 0, nr = 0, fbr = 0;
    const pihm_b200_mesh *g = nullptr;
    std::vector<int> owner_e, owner_r;              // owner rank of each global element / river
    struct Part {
        std::vector<int> elems, rivs;               // global ids (0-based): owned first, then ghosts by owner
        int nown_e = 0, nown_r = 0;
        std::vector<int> nbr;                       // neighbour ranks (sorted)
        std::vector<std::vector<int>> recv_e, recv_r;   // per neighbour: global ids received (ghost order)
        std::vector<std::vector<int>> send_e, send_r;   // per neighbour: LOCAL owned indices to send
    };
    std::vector<Part> parts;
};

extern "C" {

pihm_b200_partition *pihm_b200_partition_create(const pihm_b200_mesh *m, int nparts)
{
    if (!m || nparts < 1 || m->nelem < nparts) { pb::set_error("partition_create: bad argument"); return nullptr; }
    const int ne = m->nelem, nr = m->nriver;
    auto EI = [&](int c, int e) { return m->elem_i32[(size_t)c * ne + e]; };
    auto RI = [&](int c, int r) { return m->riv_i32[(size_t)c * nr + r]; };
    pihm_b200_partition *P = new pihm_b200_partition();
    P->nparts = nparts; P->ne = ne; P->nr = nr; P->fbr = m->fbr; P->g = m;

    // locality order, then equal contiguous chunks
    std::vector<int> nabr((size_t)3 * ne), lr((size_t)2 * std::max(nr, 1)), order(ne);
    for (int j = 0; j < 3; j++)
        for (int e = 0; e < ne; e++) nabr[(size_t)j * ne + e] = EI(PB_EI_NABR0 + j, e);
    for (int r = 0; r < nr; r++) { lr[r] = RI(PB_RI_LEFTELE, r); lr[nr + r] = RI(PB_RI_RIGHTELE, r); }
    pb::patch_order(ne, nr, nabr.data(), lr.data(), 128, order.data());
    P->owner_e.assign(ne, 0);
    std::vector<std::vector<int>> owned_e(nparts), owned_r(nparts);
    for (int k = 0; k < ne; k++) {
        const int p = (int)(((long long)k * nparts) / ne);
        P->owner_e[order[k]] = p;
        owned_e[p].push_back(order[k]);
    }
    P->owner_r.assign(nr, 0);
    for (int r = 0; r < nr; r++) {
        P->owner_r[r] = P->owner_e[RI(PB_RI_LEFTELE, r) - 1];
        owned_r[P->owner_r[r]].push_back(r);
    }
    // river adjacency helpers
    std::vector<std::vector<int>> up(nr);
    for (int r = 0; r < nr; r++) { const int d = RI(PB_RI_DOWN, r); if (d > 0) up[d - 1].push_back(r); }
    auto elem_rivers = [&](int e, std::vector<int> &out) {
        for (int j = 0; j < 3; j++) { const int n = EI(PB_EI_NABR0 + j, e); if (n < 0) out.push_back(-n - 1); }
    };

    P->parts.resize(nparts);
    // what each part needs from every other part (ghost elements / rivers, global ids)
    std::vector<std::map<int, std::vector<int>>> need_e(nparts), need_rv(nparts);
    for (int p = 0; p < nparts; p++) {
        pihm_b200_partition::Part &pt = P->parts[p];
        std::vector<char> mine(ne, 0);
        for (int e : owned_e[p]) mine[e] = 1;
        // Owned elements in the order [interior | boundary], each in locality order: an interior
        // element reads no ghost in the first RHS kernel (all three neighbours and every adjacent
        // river are owned), so that kernel works on the interior tiles while the halo records of
        // this evaluation are still on their way and waits for them only before the rest.
        std::stable_partition(owned_e[p].begin(), owned_e[p].end(), [&](int e) {
            for (int j = 0; j < 3; j++) {
                const int n = EI(PB_EI_NABR0 + j, e);
                if (n > 0 && !mine[n - 1]) return false;
                if (n < 0 && P->owner_r[-n - 1] != p) return false;
            }
            return true;
        });
        pt.elems = owned_e[p]; pt.nown_e = (int)owned_e[p].size();
        pt.rivs = owned_r[p]; pt.nown_r = (int)owned_r[p].size();
        std::set<int> G;
        std::vector<int> ring1, tmp;
        // two rings of edge neighbours
        for (int e : owned_e[p])
            for (int j = 0; j < 3; j++) {
                const int n = EI(PB_EI_NABR0 + j, e);
                if (n > 0 && !mine[n - 1] && G.insert(n - 1).second) ring1.push_back(n - 1);
            }
        for (int e : ring1)
            for (int j = 0; j < 3; j++) {
                const int n = EI(PB_EI_NABR0 + j, e);
                if (n > 0 && !mine[n - 1]) G.insert(n - 1);
            }
        // rivers computed locally: owned, touching owned elements, upstream of owned
        std::set<int> C(owned_r[p].begin(), owned_r[p].end());
        for (int e : owned_e[p]) { tmp.clear(); elem_rivers(e, tmp); for (int r : tmp) C.insert(r); }
        for (int r : owned_r[p]) for (int u : up[r]) C.insert(u);
        std::set<int> need_r = C;
        for (int r : C) { const int d = RI(PB_RI_DOWN, r); if (d > 0) need_r.insert(d - 1); }
        for (int r : need_r)
            for (int side = 0; side < 2; side++) {
                const int b = RI(side ? PB_RI_RIGHTELE : PB_RI_LEFTELE, r) - 1;
                if (!mine[b]) G.insert(b);
            }
        // rivers that only lend a stage to the friction slope of local elements
        for (int e : G) { tmp.clear(); elem_rivers(e, tmp); for (int r : tmp) need_r.insert(r); }
        for (int e : G) need_e[p][P->owner_e[e]].push_back(e);
        for (int r : need_r) if (P->owner_r[r] != p) need_rv[p][P->owner_r[r]].push_back(r);
    }
    // neighbour relation made symmetric (a part also talks to those that need it)
    for (int p = 0; p < nparts; p++) {
        std::set<int> nb;
        for (auto &kv : need_e[p]) nb.insert(kv.first);
        for (auto &kv : need_rv[p]) nb.insert(kv.first);
        for (int q = 0; q < nparts; q++)
            if (q != p && (need_e[q].count(p) || need_rv[q].count(p))) nb.insert(q);
        P->parts[p].nbr.assign(nb.begin(), nb.end());
    }
    for (int p = 0; p < nparts; p++) {
        pihm_b200_partition::Part &pt = P->parts[p];
        const int nn = (int)pt.nbr.size();
        pt.recv_e.assign(nn, {}); pt.recv_r.assign(nn, {}); pt.send_e.assign(nn, {}); pt.send_r.assign(nn, {});
        for (int k = 0; k < nn; k++) {
            const int q = pt.nbr[k];
            if (need_e[p].count(q)) pt.recv_e[k] = need_e[p][q];
            if (need_rv[p].count(q)) pt.recv_r[k] = need_rv[p][q];
            pt.elems.insert(pt.elems.end(), pt.recv_e[k].begin(), pt.recv_e[k].end());
            pt.rivs.insert(pt.rivs.end(), pt.recv_r[k].begin(), pt.recv_r[k].end());
        }
    }
    // send lists: what q receives from p, as p's LOCAL owned indices, in q's ghost order
    for (int p = 0; p < nparts; p++) {
        pihm_b200_partition::Part &pt = P->parts[p];
        std::map<int, int> loc_e, loc_r;
        for (int k = 0; k < pt.nown_e; k++) loc_e[pt.elems[k]] = k;
        for (int k = 0; k < pt.nown_r; k++) loc_r[pt.rivs[k]] = k;
        for (size_t k = 0; k < pt.nbr.size(); k++) {
            const int q = pt.nbr[k];
            if (need_e[q].count(p)) for (int e : need_e[q][p]) pt.send_e[k].push_back(loc_e.at(e));
            if (need_rv[q].count(p)) for (int r : need_rv[q][p]) pt.send_r[k].push_back(loc_r.at(r));
        }
    }
    return P;
}

void pihm_b200_partition_destroy(pihm_b200_partition *P) { delete P; }

// sizes[8] = {ne_local, nr_local, ne_owned, nr_owned, n_neighbours, n_send_elem, n_send_riv, reserved}
int pihm_b200_partition_sizes(const pihm_b200_partition *P, int part, int32_t *sizes)
{
    if (!P || part < 0 || part >= P->nparts) return -1;
    const auto &pt = P->parts[part];
    size_t se = 0, sr = 0;
    for (auto &v : pt.send_e) se += v.size();
    for (auto &v : pt.send_r) sr += v.size();
    sizes[0] = (int)pt.elems.size(); sizes[1] = (int)pt.rivs.size();
    sizes[2] = pt.nown_e; sizes[3] = pt.nown_r; sizes[4] = (int)pt.nbr.size();
    sizes[5] = (int)se; sizes[6] = (int)sr; sizes[7] = 0;
    return 0;
}

// Fill the local column tables (same layout as the global ones, local numbering)
// and the exchange maps.  Array sizes from pihm_b200_partition_sizes:
//   elem_f64 [PB_E_NCOL][ne_local], elem_i32 [PB_EI_NCOL][ne_local], riv_* likewise,
//   elem_gid [ne_local], riv_gid [nr_local]  (global 0-based ids),
//   nbr_rank [nn], send_e_ptr [nn+1], send_e_idx [n_send_elem], recv_e_cnt [nn],
//   send_r_ptr [nn+1], send_r_idx [n_send_riv], recv_r_cnt [nn]
int pihm_b200_partition_fill(const pihm_b200_partition *P, int part, double *ef, int32_t *ei, double *rf,
                             int32_t *ri, int32_t *elem_gid, int32_t *riv_gid, int32_t *nbr_rank,
                             int32_t *send_e_ptr, int32_t *send_e_idx, int32_t *recv_e_cnt,
                             int32_t *send_r_ptr, int32_t *send_r_idx, int32_t *recv_r_cnt)
{
    if (!P || part < 0 || part >= P->nparts) return -1;
    const pihm_b200_mesh *m = P->g;
    const auto &pt = P->parts[part];
    const int ne = P->ne, nr = P->nr, nl = (int)pt.elems.size(), rl = (int)pt.rivs.size();
    std::vector<int> loc_e(ne, -1), loc_r(std::max(nr, 1), -1);
    for (int k = 0; k < nl; k++) loc_e[pt.elems[k]] = k;
    for (int k = 0; k < rl; k++) loc_r[pt.rivs[k]] = k;
    for (int k = 0; k < nl; k++) {
        const int e = pt.elems[k];
        elem_gid[k] = e;
        for (int c = 0; c < PB_E_NCOL; c++) ef[(size_t)c * nl + k] = m->elem_f64[(size_t)c * ne + e];
        for (int c = 0; c < PB_EI_NCOL; c++) ei[(size_t)c * nl + k] = m->elem_i32[(size_t)c * ne + e];
        for (int j = 0; j < 3; j++) {
            const int n = m->elem_i32[(size_t)(PB_EI_NABR0 + j) * ne + e];
            int v = 0;
            if (n > 0) v = (loc_e[n - 1] >= 0) ? loc_e[n - 1] + 1 : 0;
            else if (n < 0) v = (loc_r[-n - 1] >= 0) ? -(loc_r[-n - 1] + 1) : 0;
            ei[(size_t)(PB_EI_NABR0 + j) * nl + k] = v;
            if (v == 0 && n != 0) {      // link leaves the local set: close it as a no-flow edge
                ei[(size_t)(PB_EI_BC0 + j) * nl + k] = 0;
                ei[(size_t)(PB_EI_FBRBC0 + j) * nl + k] = 0;
            }
        }
    }
    for (int k = 0; k < rl; k++) {
        const int r = pt.rivs[k];
        riv_gid[k] = r;
        for (int c = 0; c < PB_R_NCOL; c++) rf[(size_t)c * rl + k] = m->riv_f64[(size_t)c * nr + r];
        for (int c = 0; c < PB_RI_NCOL; c++) ri[(size_t)c * rl + k] = m->riv_i32[(size_t)c * nr + r];
        const int l = m->riv_i32[(size_t)PB_RI_LEFTELE * nr + r] - 1;
        const int rt = m->riv_i32[(size_t)PB_RI_RIGHTELE * nr + r] - 1;
        // stage-only ghost rivers may miss a bank locally: point it at any local element
        const int ll = (loc_e[l] >= 0) ? loc_e[l] : ((loc_e[rt] >= 0) ? loc_e[rt] : 0);
        const int lr = (loc_e[rt] >= 0) ? loc_e[rt] : ll;
        ri[(size_t)PB_RI_LEFTELE * rl + k] = ll + 1;
        ri[(size_t)PB_RI_RIGHTELE * rl + k] = lr + 1;
        const int d = m->riv_i32[(size_t)PB_RI_DOWN * nr + r];
        if (d > 0) ri[(size_t)PB_RI_DOWN * rl + k] = (loc_r[d - 1] >= 0) ? loc_r[d - 1] + 1 : -3;
    }
    int se = 0, sr = 0;
    for (size_t k = 0; k < pt.nbr.size(); k++) {
        nbr_rank[k] = pt.nbr[k];
        send_e_ptr[k] = se; send_r_ptr[k] = sr;
        for (int v : pt.send_e[k]) send_e_idx[se++] = v;
        for (int v : pt.send_r[k]) send_r_idx[sr++] = v;
        recv_e_cnt[k] = (int)pt.recv_e[k].size();
        recv_r_cnt[k] = (int)pt.recv_r[k].size();
    }
    send_e_ptr[pt.nbr.size()] = se; send_r_ptr[pt.nbr.size()] = sr;
    return 0;
}

}  // extern "C"
