function [y, u] = ClosedLoopNMPC(x0_model, x_control, u0, r, N, Nu, Q, W, nit, ub1, lb1, inK, Ts, noise)
% CLOSEDLOOPNMPC  Drop-in replacement of "Explicit NMPC/ClosedLoopNMPC.m":1 on libmpcgpu.so (kernel k_ssnmpc).
%
%   [y, u] = ClosedLoopNMPC(x0_model, x_control, u0, r, N, Nu, Q, W, nit, ub1, lb1, inK, Ts)
%
% Same argument list and outputs (signals x time) as the reference; main.m:65 runs unchanged.  The controller call is
% NMPC_Controller.m:1 (single shooting, per-input control horizons, offsets from u(k-1), the model-deviation term of
% :106-123), the model plant_model.m:1-56.  Differences, stated: RK4 with 4 sub-steps per sample instead of ode23t / ode45;
% Gauss-Newton with an exact box-QP step instead of fmincon-SQP (same minimiser); and the state noise the reference draws
% inside its loop (0.01*randn per sample, ClosedLoopNMPC.m:76,89) is the optional 14th argument -- pass
% 0.01*randn(numel(x0_model), nit) to reproduce a noisy run, omit it for the noise-free one.
%
% A handle is created per call (a closed loop of 150 samples runs in a few milliseconds on the device); for sweeps over
% (N, Nu, Q, W) keep one handle and call mpcgpu_mex('ssnmpc_eval', hs, N, Nu, Q, W) with n x 1 / n x 2 arrays instead.
Ps = struct('nit', nit, 'pmax', max(31, N(1)), 'inK', inK, 'Ts', Ts, 'x_control', x_control(:)', ...
            'x0', x0_model(:)', 'u0', u0(:)', 'lb', lb1(:)', 'ub', ub1(:)', 'r', reshape(r(:, 1:nit)', 1, []));
hs = mpcgpu_mex('ssnmpc_create', Ps);
cleanup = onCleanup(@() mpcgpu_mex('ssnmpc_destroy', hs));
if nargin < 14
    [y, u] = mpcgpu_mex('ssnmpc_closedloop', hs, r, N, Nu, Q, W);
else
    [y, u] = mpcgpu_mex('ssnmpc_closedloop', hs, r, N, Nu, Q, W, noise);
end
end
