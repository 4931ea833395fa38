// plan.h -- on-disk layout of a .b200w plan (the serialized network: op list + packed weights).
// Written by vehicle-cv-adas_b200/plan.py, read by engine.cu.  Plays the role of the reference's
// serialized TensorRT engine / ONNX file (coreEngine.py:54-55,164-166); the packer replaces
// convertOnnxToTensorRT.py / onnxQuantization.py for this runtime.
#pragma once
#include <stdint.h>

namespace adas {

static const char kPlanMagic[8] = {'B', '2', '0', '0', 'P', 'L', 'A', 'N'};
static const uint32_t kPlanVersion = 1;

#pragma pack(push, 1)
struct PlanHeader {
    char magic[8];
    uint32_t version;
    uint32_t model_kind;          // ADAS_MODEL_*
    uint32_t in_c, in_h, in_w;    // network input binding (NCHW semantic)
    uint32_t n_buffers, n_ops, n_tensors, n_outputs;
    uint32_t meta[16];            // YOLO: [0]=nc [1]=n_anchors(total)   UFLD: [0]=ngr [1]=ncr [2]=ngc [3]=ncc [4]=nl [5]=total_dim
    uint64_t blob_offset, blob_bytes;
};
struct PlanBuffer {               // activation buffer: [batch * rows_per_img, C] elements
    uint32_t rows_per_img;
    uint32_t C;                   // row stride in elements
    uint32_t dtype;               // 0 = fp16, 1 = fp32
    uint32_t H, W;                // > 0: padded-NHWC geometry, rows_per_img == (H+2)*(W+2)
    uint32_t flags;
};
struct PlanOp {
    uint32_t type;                // PlanOpType
    int32_t p[23];
    float f[4];
};
struct PlanTensor {
    uint64_t offset, bytes;       // relative to blob_offset
    uint32_t dtype;               // 0 = fp16, 1 = fp32
    uint32_t pad;
};
struct PlanOutput {
    uint32_t buffer, coff, C, stride;   // stride: YOLO level stride (8/16/32)
};
#pragma pack(pop)

enum PlanOpType : uint32_t {
    OP_GEMM = 1,       // p: a_buf a_coff Kc ntaps w_tensor bias_tensor N act res_buf res_coff res_pre_act out_buf out_coff masked transposed BN s2 MT
                       //    (BN / MT > 0 force the tile shape: test hooks, 0 = cost model + autotune)
    OP_IM2COL = 2,     // p: in_buf in_coff Cin kh kw stride pad out_buf
    OP_MAXPOOL = 3,    // p: in_buf in_coff C k s pad out_buf out_coff
    OP_UPSAMPLE2X = 4, // p: in_buf in_coff C out_buf out_coff
    OP_STEMPACK = 6,   // p: in_buf(image, C=4) out_buf : 7x7 stride-2 stem re-layout, see elementwise.cu stempack_kernel
    OP_STEMCONV = 7,   // p: in_buf(image, C=4) w_tensor bias_tensor Cout k pad act out_buf out_coff : k x k stride-2 stem conv, stem_conv.cu
                       //    (weights packed [Cout][k][round_up(4k,16)])
    OP_LAYERNORM = 5,  // p: in_buf d_len gamma_tensor beta_tensor out_buf d_norm ; f0 = eps (statistics over d_norm entries; the
                       //    other d_len - d_norm slab entries are structural zeros with gamma = beta = 0)
};

}  // namespace adas
