-- optim.npe_adam / optim.npe_rmsprop: same signature as optim.adam / optim.rmsprop (opfunc, x, state) -> x, {fx}
-- (reference call site main.lua:301-302; state.learningRate is mutated from outside, main.lua:384).  The update runs
-- on the device vector behind model.network; x (the host mirror) is refreshed afterwards so `new_params ==
-- model.theta.params` (main.lua:304) keeps holding.
require 'optim'
local npeffi = require 'npe_ffi'
local C = npeffi.C

local function step(kind, opfunc, x, state, network)
    local fx, dfdx = opfunc(x)                       -- feval_train: fp + bp (gradient now on the device AND in dfdx)
    local lr = state.learningRate or 1e-3
    local h = network.handle
    if kind == 'adam' then
        npeffi.check(h, C.npe_adam_step(h, lr, state.beta1 or 0.9, state.beta2 or 0.999, state.epsilon or 1e-8))
    else
        npeffi.check(h, C.npe_rmsprop_step(h, lr, state.alpha or 0.99, state.epsilon or 1e-8))
    end
    network:pull()                                   -- x aliases network.weight (getParameters): refreshed in place
    return x, {fx}
end

function optim.npe_adam(opfunc, x, state) return step('adam', opfunc, x, state, state.network) end
function optim.npe_rmsprop(opfunc, x, state) return step('rmsprop', opfunc, x, state, state.network) end
