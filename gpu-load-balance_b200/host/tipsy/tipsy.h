// tipsy.h — reader/writer for tipsy particle snapshots, the input format the reference's Init service
// was meant to read (andrinr/gpu-load-balance src/services/init.cu:5,54-59: `TipsyIO io; io.open(path);
// io.count(); io.load(particles)`; the reference's own src/tipsy/ is not in its repository, CMakeLists.txt:74).
//
// File layout (tipsy "dump"): header {double time; int nbodies, ndim, nsph, ndark, nstar; [int pad]},
// then nsph gas records (12 floats: mass pos[3] vel[3] rho temp hsmooth metals phi), ndark dark records
// (9 floats: mass pos[3] vel[3] eps phi), nstar star records (11 floats: mass pos[3] vel[3] metals tform
// eps phi).  "Standard" files (.std) are big-endian (XDR) with a 32-byte header; "native" files are
// host-endian with a 28- or 32-byte header.  The variant is detected from ndim and the file size.
//
// Only positions are used by ORB; load() de-interleaves them into the x / y / z columns the device
// context takes (orb_upload_xyz), for any contiguous slice of bodies so every rank reads just its shard.
#ifndef ORB_HOST_TIPSY_H
#define ORB_HOST_TIPSY_H

#include <cstdint>
#include <cstdio>
#include <string>

class TipsyIO {
public:
    TipsyIO() = default;
    ~TipsyIO() { close(); }
    TipsyIO(const TipsyIO &) = delete;
    TipsyIO &operator=(const TipsyIO &) = delete;

    // false on failure; error() says why
    bool open(const char *path);
    void close();
    // bodies in the file (init.cu:57 prints it next to the requested count)
    uint64_t count() const { return nBodies_; }
    uint64_t nGas() const { return nSph_; }
    uint64_t nDark() const { return nDark_; }
    uint64_t nStar() const { return nStar_; }
    double time() const { return time_; }
    bool standard() const { return bigEndian_; }
    int headerBytes() const { return headerBytes_; }
    // positions of bodies [first, first + n) in file order (gas, dark, star)
    bool load(uint64_t first, uint64_t n, float *x, float *y, float *z);
    const std::string &error() const { return error_; }

    // all bodies written as dark particles (mass 1/n, zero velocity); standard = big-endian .std
    static bool writePositions(const char *path, uint64_t n, const float *x, const float *y, const float *z,
                               bool standard, std::string *error);

private:
    bool fail(const std::string &why) { error_ = why; close(); return false; }
    FILE *file_ = nullptr;
    std::string error_;
    double time_ = 0.0;
    uint64_t nBodies_ = 0, nSph_ = 0, nDark_ = 0, nStar_ = 0;
    int headerBytes_ = 0;
    bool bigEndian_ = false;
};

#endif
