// mpcv_model_select.h — ModelT of this translation unit from -DMPCV_INST_MODEL=<MPCV_MODEL_* id>
#pragma once
#if MPCV_INST_MODEL == 0
using ModelT = mpcv::Unicycle<0>;
#elif MPCV_INST_MODEL == 1
using ModelT = mpcv::Unicycle<1>;
#elif MPCV_INST_MODEL == 2
using ModelT = mpcv::Unicycle<2>;
#elif MPCV_INST_MODEL == 3
using ModelT = mpcv::Linear<3, false>;
#elif MPCV_INST_MODEL == 4
using ModelT = mpcv::Linear<4, false>;
#elif MPCV_INST_MODEL == 5
using ModelT = mpcv::Linear<4, true>;
#elif MPCV_INST_MODEL == 6
using ModelT = mpcv::Linear<3, true>;
#elif MPCV_INST_MODEL == 7
using ModelT = mpcv::FrenetBicycle;
#else
#error "unknown MPCV_INST_MODEL"
#endif
static_assert(ModelT::MODEL_ID == MPCV_INST_MODEL, "model id mismatch");
