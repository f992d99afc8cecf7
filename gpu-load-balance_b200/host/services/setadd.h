// setadd.h — builds the PST over the rank threads (reference: src/services/setadd.h, setadd.cpp:16-45).
#ifndef ORB_HOST_SETADD_H
#define ORB_HOST_SETADD_H
#include "pst.h"

class ServiceSetAdd : public mdl::BasicService {
    PST node_pst;
public:
    struct input {
        int idLower;
        int idUpper;
        input() = default;
        input(int idUpper_) : idLower(0), idUpper(idUpper_) {}
        input(int idLower_, int idUpper_) : idLower(idLower_), idUpper(idUpper_) {}
    };
    typedef void output;
    explicit ServiceSetAdd(PST pst) : BasicService(PST_SETADD, sizeof(input), "SetAdd"), node_pst(pst) {}

protected:
    virtual int operator()(int nIn, void *pIn, void *pOut) override;
    void SetAdd(PST pst, input *in);
};
#endif
