function [y,u,t] = mpcgpu_validation_run(h, real_plant, Mgain, hl, r, v, N, Nu, delta, lambda, nit)
% MPCGPU_VALIDATION_RUN  The validation run of the case scripts against a plant that differs from the prediction model
% (MPC-Tuning/Shell3x3.m:271-286, WoodBerry.m:263-278, Shell7x5.m:293-306):
%     options = mpcsimopt(mpc_toolbox); options.Model = plant; [y,t,u] = sim(mpc_toolbox,nit,r,[],options);
% on libmpcgpu.so.  h: handle of mpcgpu_mex('create', P); real_plant: the scaled DISCRETE real process (c2d(L*Psr*R,Ts), first
% order plus dead time channels like the model's); Mgain: the controller's estimator gain in the library's state order
% (include/mpcgpu.h, mpcgpu_set_mismatch: channel states, MV delay-line states w_j(k-1-q) q = 0..hl-1, output-disturbance
% states), hl: delay-line states per MV.  Outputs as closedloop_toolbox: signals x time.
[num, den] = tfdata(real_plant);
[ny, nw] = size(num);
Pl.a = zeros(nw, ny); Pl.b0 = Pl.a; Pl.b1 = Pl.a;
for i = 1:ny
    for j = 1:nw
        dd = den{i,j}; nn = num{i,j};
        Pl.a(j,i) = -dd(2) / dd(1); Pl.b0(j,i) = nn(1) / dd(1); Pl.b1(j,i) = nn(2) / dd(1);
    end
end
Pl.d = int32(real_plant.IODelay)';
mpcgpu_mex('mismatch', h, Pl, Mgain, hl);
cleanup = onCleanup(@() mpcgpu_mex('mismatch', h));          % back to the nominal evaluation whatever happens
[y,u,t] = mpcgpu_mex('closedloop', h, r, v, N, Nu, delta, lambda, nit);
end
