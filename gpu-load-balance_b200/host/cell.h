// cell.h — ORB tree node, wire-compatible with the reference's `struct Cell`
// (andrinr/gpu-load-balance src/cell.h:9-136): same field order and 52-byte layout (it is the
// input type of every service), same member names so master() reads the same.  The arithmetic
// that matters for parity is kept literal: getCut() adds in float and halves in double
// (cell.h:74-76); setCutAxis() uses a strict '>' starting from 0 so the lowest axis wins ties and
// an all-zero box yields -1 (cell.h:102-121).
#ifndef ORB_HOST_CELL_H
#define ORB_HOST_CELL_H

#include <cstdio>
#include <tuple>

struct Cell {
    int id;
    int nLeafCells;
    int prevCutAxis;
    int cutAxis;
    bool foundCut;
    float cutMarginLeft;
    float cutMarginRight;
    float lower[3], upper[3];

    Cell() = default;   // trivial: the struct travels through services by memcpy

    Cell(int id_, int nLeafCells_, const float *lower_, const float *upper_)
        : id(id_), nLeafCells(nLeafCells_), prevCutAxis(-1), cutAxis(-1), foundCut(false),
          cutMarginLeft(0.0f), cutMarginRight(0.0f) {
        for (int k = 0; k < 3; ++k) {
            lower[k] = lower_[k];
            upper[k] = upper_[k];
        }
    }

    // implicit binary heap, root id 0 (cell.h:49-59)
    int getLeftChildId() const { return 2 * id + 1; }
    int getRightChildId() const { return 2 * id + 2; }
    int getParentId() const { return (id + 1) / 2 - 1; }
    int getTotalNumberOfCells() const { return 2 * nLeafCells - 1; }

    // ceil(log2(nLeafCells)) (cell.h:65-67) in integers: smallest k with 2^k >= nLeafCells
    int getNLevels() const {
        int k = 0;
        while ((1LL << k) < (long long)nLeafCells) ++k;
        return k;
    }
    int getNCellsOnLastLevel() const { return 2 * nLeafCells - (1 << getNLevels()); }   // cell.h:69-72

    float getCut() const { return (cutMarginRight + cutMarginLeft) / 2.0; }             // cell.h:74-76

    // cell.h:78-100: left child takes ceil(n/2) leaf cells; boxes are cut at getCut() on cutAxis
    std::tuple<Cell, Cell> cut() const {
        const int nLeft = (nLeafCells + 1) / 2;
        const float c = getCut();
        Cell l(getLeftChildId(), nLeft, lower, upper);
        Cell r(getRightChildId(), nLeafCells - nLeft, lower, upper);
        l.upper[cutAxis] = c;
        r.lower[cutAxis] = c;
        l.prevCutAxis = r.prevCutAxis = cutAxis;
        return std::make_tuple(l, r);
    }

    // longest geometric side (cell.h:102-121)
    void setCutAxis() {
        int best = -1;
        float bestSize = 0.0f;
        for (int d = 0; d < 3; ++d) {
            const float size = upper[d] - lower[d];
            if (size > bestSize) {
                bestSize = size;
                best = d;
            }
        }
        cutAxis = best;
    }

    // bisection starts from the box faces on the cut axis (cell.h:123-126)
    void setCutMargin() {
        cutMarginLeft = lower[cutAxis];
        cutMarginRight = upper[cutAxis];
    }

    void log() const {
        std::printf("cell %d: children %d/%d parent %d leaves %d axis %d (prev %d) found %d\n", id, getLeftChildId(),
                    getRightChildId(), getParentId(), nLeafCells, cutAxis, prevCutAxis, (int)foundCut);
        std::printf("  box [%f %f %f] .. [%f %f %f]  margins %f .. %f\n", lower[0], lower[1], lower[2], upper[0],
                    upper[1], upper[2], cutMarginLeft, cutMarginRight);
    }
};

static_assert(sizeof(Cell) == 52, "Cell must keep the reference's 52-byte wire layout");

#endif
