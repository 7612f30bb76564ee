function P = mpcgpu_problem(mpcobj, Pze, Xsp, mdv, Yref, dmin, nit, nbp, nbc)
% MPCGPU_PROBLEM  The struct mpcgpu_mex('create', P) takes, built from what MPCTuning.m has in hand after its scaling block
% (MPC-Tuning/MPC_Tuning/MPCTuning.m:154-262): the scaled `mpc` object, its scaled discrete plant Pze = L*Pz*R (first-order-
% plus-dead-time channels, MVs then MDs), the scaled set-point Xsp (nit x ny or ny x nit), measured disturbance mdv
% (nit x nd, may be empty), reference trajectory Yref (ny x nit or nit x ny), the per-output minimum dead times dmin
% (MPCTuning.m:257-262) and the sizes of the binary horizon codes nbp / nbc (prediction / control horizon bits).
%
%   P = mpcgpu_problem(mpcobj, Pze, Xsp, mdv, Yref, dmin, nit, nbp, nbc);   Par.gpu = mpcgpu_mex('create', P);
%
% Layout notes: the C ABI is row-major; a MATLAB ny x nw matrix is handed over transposed (nw x ny column-major has the same
% bytes).  Signals: r nit x ny and v nit x nd time-major, yref ny x nit.
[num, den] = tfdata(Pze);
[ny, nw] = size(num);
nu = numel(mpcobj.MV); nd = nw - nu;
a = zeros(ny, nw); b0 = a; b1 = a;
for i = 1:ny
    for j = 1:nw
        dd = den{i,j}; nn = num{i,j};
        if numel(dd) ~= 2 || numel(nn) ~= 2, error('mpcgpu:arg', 'channel (%d,%d) is not first order plus dead time', i, j); end
        a(i,j) = -dd(2) / dd(1); b0(i,j) = nn(1) / dd(1); b1(i,j) = nn(2) / dd(1);
    end
end
P = struct();
P.ny = ny; P.nu = nu; P.nd = nd; P.nit = nit; P.pmax = 2^nbp - 1; P.mmax = 2^nbc - 1; P.inK = 10;   % inK: VNS2.m:43
P.Ts = Pze.Ts;
P.a = a'; P.b0 = b0'; P.b1 = b1'; P.d = int32(Pze.IODelay)';
P.umin = [mpcobj.MV.Min]; P.umax = [mpcobj.MV.Max]; P.dumin = [mpcobj.MV.RateMin]; P.dumax = [mpcobj.MV.RateMax];
P.ymin = [mpcobj.OV.Min]; P.ymax = [mpcobj.OV.Max]; P.ecr_min = [mpcobj.OV.MinECR]; P.ecr_max = [mpcobj.OV.MaxECR];
P.su = [mpcobj.MV.ScaleFactor]; P.sy = [mpcobj.OV.ScaleFactor]; P.rho_ecr = mpcobj.Weights.ECR;
if size(Xsp,1) ~= nit, Xsp = Xsp'; end            % row2col.m
P.r = Xsp';                                       % ny x nit column-major == nit x ny row-major
if nd > 0
    if size(mdv,1) ~= nit, mdv = mdv'; end
    P.v = mdv';
end
if size(Yref,1) == nit && size(Yref,2) == ny, Yref = Yref'; end
P.yref = Yref';                                   % nit x ny column-major == ny x nit row-major
P.dmin = int32(dmin(:))';
end
