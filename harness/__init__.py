"""Measurement harnesses around the hot path (BASELINE.json configs[2] and configs[4]).

NOT part of the product package: a stock-PyTorch stand-in for the rest of the reference's
ParkingModel (camera encoder, BEV encoder, fusion transformer, control decoder, segmentation
head, losses) so that the lift-splat library can be timed inside a full training step under
DDP and inside the closed-loop agent's predict().  Those sub-networks are out of scope for the
sm_100a work (BASELINE.json north_star: they "stay in stock PyTorch").
"""
