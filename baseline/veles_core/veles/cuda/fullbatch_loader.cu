// Core kernel of the (absent) Veles platform, written for the reference-arm shim: gathers the
// minibatch rows (and labels) by index out of the device-resident dataset.
extern "C"
__global__ void fill_minibatch_data_labels(const original_data_dtype *original_data,
                                           minibatch_data_dtype *minibatch_data,
                                           const int count, const int sample_size,
                                           const int *original_labels, int *minibatch_labels,
                                           const int *indices) {
  size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t total = (size_t)MAX_MINIBATCH_SIZE * SAMPLE_SIZE;
  if (idx >= total) return;
  int sample = (int)(idx / SAMPLE_SIZE);
  int offs = (int)(idx % SAMPLE_SIZE);
  if (sample < count) {
    int src = indices[sample];
    minibatch_data[idx] = (minibatch_data_dtype)original_data[(size_t)src * SAMPLE_SIZE + offs];
    if (!offs && original_labels) minibatch_labels[sample] = original_labels[src];
  } else {
    minibatch_data[idx] = 0;
    if (!offs && original_labels) minibatch_labels[sample] = -1;
  }
}
