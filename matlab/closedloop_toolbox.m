function [y,u,t,ys,uopt] = closedloop_toolbox(mpc_toolbox,r,v,N,Nu,delta,lambda,nit)
% CLOSEDLOOP_TOOLBOX  Drop-in replacement of MPC-Tuning/MPC_Tuning/closedloop_toolbox.m:1 on libmpcgpu.so (B200).
%
%   [y,u,t,ys,uopt] = closedloop_toolbox(mpc_toolbox,r,v,N,Nu,delta,lambda,nit)
%
% Same argument list, output count and orientations as the reference (signals x time after col2row,
% closedloop_toolbox.m:103-107; r and v in either orientation, row2col.m:3-8; N and Nu may be vectors, their max is
% used, :38-40).  The only difference for a caller: mpc_toolbox is the uint64 handle returned by
%   Par.gpu = mpcgpu_mex('create', P)        % P: scaled plant, limits, scale factors, Par.Xsp, Par.Yref (MPCTuning.m:154-340)
% instead of the scaled `mpc` object (GAM_fun.m:81, VNS2.m:153/168, Shell3x3.m:231 pass Par.mpcobj there: pass Par.gpu).
% A failed simulation raises mpcgpu:candidate, which the reference's try/catch blocks (GAM_fun.m:80-91,
% VNS2.m:151-163) print and skip exactly as they do a Toolbox exception.  The handle's own signals are not modified.
[y,u,t,ys,uopt] = mpcgpu_mex('closedloop', mpc_toolbox, r, v, N, Nu, delta, lambda, nit);
end
