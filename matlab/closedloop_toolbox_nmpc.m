function [y,u,yopt,uopt] = closedloop_toolbox_nmpc(nmpcobj,model,init,r,N,Nu,delta,lambda,nit)
% CLOSEDLOOP_TOOLBOX_NMPC  Drop-in replacement of MPC-Tuning/MPC_Tuning/closedloop_toolbox_nmpc.m:1 on libmpcgpu.so.
%
%   [y,u,yopt,uopt] = closedloop_toolbox_nmpc(nmpcobj,model,init,r,N,Nu,delta,lambda,nit)
%
% Same argument list and outputs (signals x time) as the reference.  nmpcobj is the uint64 handle returned by
%   Par.gpu = mpcgpu_mex('nmpc_create', Pn)  % Pn: x0, u0, bounds, scale factors, Ts, r, yref (VanDeVusse_NMPC.m:35-204)
% `model` and `init` are accepted for signature compatibility and ignored: the plant (vandevusse_model.m:39-77), the
% initial state init.x0 and input init.u0 are part of Pn.  Outputs are states 2..nx, as closedloop_toolbox_nmpc.m:73.
[y,u,yopt,uopt] = mpcgpu_mex('nmpc_closedloop', nmpcobj, r, N, Nu, delta, lambda, nit); %#ok<INUSL>
end
