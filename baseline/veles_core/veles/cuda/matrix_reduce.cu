// Core include of the (absent) Veles platform, written for the reference-arm shim.
// Included INSIDE a kernel body. Contract (see /root/reference/tests/unit/cuda/
// test_matrix_reduce.cu and cuda/weights_ortho.cu): with A, A_WIDTH, A_HEIGHT, REDUCE_SIZE and
// optionally A_COL defined, block bx reduces column bx (A_COL) or row bx of the row-major
// matrix A with REDUCE_SIZE threads; afterwards thread tx == 0 holds the total as sum + AS[0].
  __shared__ dtype AS[REDUCE_SIZE];

  const int bx = blockIdx.x;
  const int tx = threadIdx.x;

  dtype sum = 0;

#ifdef A_COL
  {
    int offs = bx + tx * A_WIDTH;
    for (int i = tx; i < A_HEIGHT; i += REDUCE_SIZE, offs += A_WIDTH * REDUCE_SIZE) {
      sum += A[offs];
    }
  }
#else
  {
    size_t offs = (size_t)bx * A_WIDTH + tx;
    for (int i = tx; i < A_WIDTH; i += REDUCE_SIZE, offs += REDUCE_SIZE) {
      sum += A[offs];
    }
  }
#endif

  AS[tx] = sum;
  __syncthreads();
  // thread 0 keeps its own partial in ``sum`` and gathers the others into AS[0]
  if (!tx) {
    dtype others = 0;
    for (int i = 1; i < REDUCE_SIZE; i++) {
      others += AS[i];
    }
    AS[0] = others;
  }
  __syncthreads();
