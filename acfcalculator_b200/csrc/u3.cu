__global__ void k(double* out, const double* __restrict__ in, int n){
  extern __shared__ double sh[];
  for (int i=threadIdx.x;i<n;i+=blockDim.x) sh[i]=in[i];
  __syncthreads();
  double acc[8]; double w[8];
  for(int j=0;j<8;++j){acc[j]=0; w[j]=in[threadIdx.x*8+j];}
  for (int i=0;i<n;++i){
    double a = sh[i];
#pragma unroll
    for(int j=0;j<8;++j) acc[j]=fma(a,w[j],acc[j]);
  }
  double s=0; for(int j=0;j<8;++j) s+=acc[j]; out[threadIdx.x]=s;
}
