function [g,status] = GAM_fun_batch(X,Par)
% GAM_FUN_BATCH  GAM_fun.m:54-115 for a whole population: X is n x (ny+nu) = [delta lambda] rows, g is n x ny.
% The batched form the reference cannot express (one closed loop per row, all rows at once on the GPU; with
% Par.gpu_multi, over all GPUs of the box).  |X| is taken and band outputs keep delta = 0 as in GAM_fun.m:56-66.
ny = Par.ny; nu = Par.nu; n = size(X,1);
delta = abs(X(:,1:ny)); lambda = abs(X(:,ny+1:ny+nu));
delta(:, Par.delta0 == 0) = 0;
N = repmat(max(Par.N), n, 1); Nu = repmat(max(Par.Nu), n, 1);
if isfield(Par,'gpu_multi')
    [g,status] = mpcgpu_mex('eval_multi', Par.gpu_multi, N, Nu, delta, lambda, 'gam');
else
    [g,status] = mpcgpu_mex('eval', Par.gpu, N, Nu, delta, lambda, 'gam');
end
end
