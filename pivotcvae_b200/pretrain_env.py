"""Response-model pre-training — drop-in for the reference's pretrain_env.py:25-139.

train_response_model keeps the reference's signature and loop: Adam(lr, weight_decay=decay), BCELoss on the sigmoid
of the slate logits, validation every epoch, save-best whole-model pickle, early termination after 3 epochs
without a 1e-4 improvement.  Forward, loss and backward run in libpcv_b200 (gather + normalise with a backward,
fused MLP block, tcgen05 backward GEMMs, one-kernel BCE).  The final "move the best model to the CPU and re-save"
step of the reference (:133-137) is left out: this library has no CPU path.
"""
import numpy as np
import torch

from .autograd import BCESigmoidFn
from .env.response_model import UserResponseModel_MLP


def _batch(batchData, device):
    slates = torch.as_tensor(np.asarray(batchData["slates"]), dtype=torch.int64, device=device)
    users = torch.as_tensor(np.asarray(batchData["users"]), dtype=torch.int64, device=device)
    targets = torch.as_tensor(np.asarray(batchData["responses"]), dtype=torch.float32, device=device)
    return slates, users, targets


def response_loss(model, batchData):
    """BCE(sigmoid(model(slates, users)), responses) -> scalar tensor (autograd attached in grad mode)."""
    dev = model.docEmbed.weight.device
    slates, users, targets = _batch(batchData, dev)
    pred = model.forward(slates, users)
    return BCESigmoidFn.apply(pred.reshape(-1), targets.reshape(-1))


def train_response_model(trainset, valset, f_size, s_size, struct, bs, epochs, lr, decay, device, model_path, logger):
    from torch.utils.data import DataLoader
    logger.log("Train user response model as simulator")
    for k, v in (("feature size", f_size), ("slate size", s_size), ("struct", struct), ("batch size", bs),
                 ("number of epoch", epochs), ("learning rate", lr), ("device", device)):
        logger.log("\t%s: %s" % (k, v))
    model = UserResponseModel_MLP(trainset.max_iid, trainset.max_uid, f_size, s_size, struct, device, trainset.noUser)
    model.to(device)
    model.differentiable = True       # forward() builds the autograd graph, embedding tables included
    trainLoader = DataLoader(trainset, batch_size=bs, shuffle=True, num_workers=0)
    valLoader = DataLoader(valset, batch_size=bs, shuffle=False, num_workers=0)
    optimizer = torch.optim.Adam(model.parameters(), lr=lr, weight_decay=decay)
    trainHistory, valHistory = [], []
    bestValLoss = float("inf")
    temper = 3
    for epoch in range(epochs):
        logger.log("Epoch " + str(epoch + 1))
        losses = []
        for batchData in trainLoader:
            optimizer.zero_grad()
            loss = response_loss(model, batchData)
            losses.append(loss.detach())
            loss.backward()
            optimizer.step()
        trainHistory.append(float(torch.stack(losses).mean()))
        logger.log("train loss: " + str(trainHistory[-1]))
        vl = []
        with torch.no_grad():
            for batchData in valLoader:
                vl.append(response_loss(model, batchData))
        valHistory.append(float(torch.stack(vl).mean()))
        logger.log("Validation Loss: " + str(valHistory[-1]))
        if epoch == 0 or valHistory[-1] < bestValLoss - 1e-4:
            torch.save(model, open(model_path, "wb"))
            logger.log("Save best model")
            temper = 3
            bestValLoss = valHistory[-1]
        else:
            temper -= 1
            logger.log("Temper down to " + str(temper))
            if temper == 0:
                logger.log("Out of temper, early termination.")
                break
    return trainHistory, valHistory
