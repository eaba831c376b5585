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

-- Fast path for train() with -rs (main.lua:299-306): one call per iteration uploads the sampled batch and enqueues
-- forward + backward + optimiser (npe_train_batch_async); the loss of iteration t-1 is read after iteration t has been
-- enqueued, so the GPU never waits for the host.  `batch` is the loader's tuple {this, context, y, ...}; returns the
-- loss of the PREVIOUS iteration (nil on the first call).  Call optim.npe_flush(state) before reading theta.params.
function optim.npe_train_async(batch, state, kind)
    local h = state.network.handle
    local this, context, y = batch[1]:contiguous(), batch[2]:contiguous(), batch[3]:contiguous()
    local B, Cn = this:size(1), context:size(2)
    state.t = (state.t or 0) + 1
    npeffi.check(h, C.npe_train_batch_async(h, npeffi.fptr(this), npeffi.fptr(context), npeffi.fptr(y), B, Cn,
                                            kind == 'rmsprop' and 1 or 0, state.learningRate or 3e-4, B, state.t % 64))
    local prev
    if state.t > 1 then
        state.loss3 = state.loss3 or torch.FloatTensor(3)
        npeffi.check(h, C.npe_loss_wait(h, (state.t - 1) % 64, npeffi.fptr(state.loss3)))
        prev = state.loss3[1]
    end
    return prev
end

function optim.npe_flush(state)
    local h = state.network.handle
    state.loss3 = state.loss3 or torch.FloatTensor(3)
    if (state.t or 0) > 0 then npeffi.check(h, C.npe_loss_wait(h, state.t % 64, npeffi.fptr(state.loss3))) end
    state.network:pull()
    return state.loss3[1]
end
