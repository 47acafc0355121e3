// A minimal operator plugin, shaped like the reference's docs/examples/custom_operations/custom_operator/customdummy/dummy.{h,cc,cu}:
// it copies its input to its output on the workspace stream.  Written against dali_b200/host/dali.h (the same names as the reference's
// operator.h) and loaded through plugin_manager.load_library; the static DALI_SCHEMA / DALI_REGISTER_OPERATOR objects register it.
#include <cuda_runtime_api.h>
#include "dali.h"

namespace other_ns {

class CustomDummy : public ::dali::Operator<::dali::GPUBackend> {
 public:
  explicit CustomDummy(const ::dali::OpSpec &spec) : ::dali::Operator<::dali::GPUBackend>(spec), scale_(spec.GetArgument<int>("repeat")) {}

 protected:
  bool SetupImpl(std::vector<::dali::OutputDesc> &output_desc, const ::dali::Workspace &ws) override {
    const auto &input = ws.Input<::dali::GPUBackend>(0);
    output_desc.resize(1);
    output_desc[0].shape = input.shape();
    output_desc[0].type = input.type();
    return true;
  }
  void RunImpl(::dali::Workspace &ws) override {
    const auto &input = ws.Input<::dali::GPUBackend>(0);
    auto &output = ws.Output<::dali::GPUBackend>(0);
    output.SetLayout(input.GetLayout());
    for (int i = 0; i < input.num_samples(); i++)
      cudaMemcpyAsync(output.raw_mutable_tensor(i), input.raw_tensor(i), input.shape().tensor_size(i) * ::dali::TypeSize(input.type()),
                      cudaMemcpyDeviceToDevice, ws.stream());
  }
 private:
  int scale_;
};

}  // namespace other_ns

DALI_SCHEMA(CustomDummy).DocStr("Make a copy of the input tensor").NumInput(1).NumOutput(1).AddOptionalArg("repeat", "unused", 1);
DALI_REGISTER_OPERATOR(CustomDummy, ::other_ns::CustomDummy, GPU);
