#!/bin/bash
